/* lys_save.c -- the reference's demo-save host (Rust: demo-save/src/{main,wrapper,ffi}.rs) in C, because this image has
 * no Rust toolchain.  Same call sequence, same defaults:
 *
 *   Fut::init (wrapper.rs:34-75): futhark_context_config_new / futhark_context_new, ljus::load -> load_obj_data,
 *       futhark_new_f32_3d / u32_1d / f32_2d / f32_1d, futhark_entry_init(seed 0, h, w, cam_conf_id 2 = LIDAR, ...,
 *       pitch, yaw, origin) -- note height before width (ffi.rs:53-66).
 *   sample_points (wrapper.rs:77-101): futhark_entry_sample_points_n(state, spp) -> new state + [h][w][4] f32,
 *       the old state is freed, futhark_values_f32_3d, the positions are kept.
 *   main (main.rs:11-32): 640x480, camera (0, 0.8, 1.8), pitch 0, yaw 0, assets/SpectrumSphere.obj, 100 samples per
 *       pixel, every pixel's point written to dump.pcd (ASCII).
 *   the commented-out image capture (main.rs:34-49): futhark_entry_sample_n_frames(state, 100) -> 8-bit RGB; enabled
 *       here with -i (the camera preset is then 0 = visual unless -c says otherwise).
 * Unlike the Rust wrapper, return codes are checked (it ignores them).  Links against the static libtracer.a, as
 * ffi.rs:1 (`#[link(name = "tracer", kind = "static")]`) does.
 *
 *   lys_save [-o scene.obj] [-w 640] [-h 480] [-s 100] [-p dump.pcd] [-i image.ppm] [-c cam_conf_id] [-d device]
 */
#define _POSIX_C_SOURCE 200809L
#include "tracer.h"
#include "lys_pcd.h"

void load_obj_data(char *obj_path, size_t *num_tris, size_t *num_mat_components, float **tri_data, uint32_t **tri_mats, float **mat_data);
void free_obj_data(float *tri_data, uint32_t *tri_mats, float *mat_data);

static void check(struct futhark_context *ctx, int res, const char *what) {
    if (res != 0) {
        char *msg = futhark_context_get_error(ctx);
        fprintf(stderr, "lys_save: %s failed (%d): %s\n", what, res, msg ? msg : "");
        free(msg);
        exit(EXIT_FAILURE);
    }
}

int main(int argc, char **argv) {
    const char *obj = "assets/SpectrumSphere.obj", *pcd = "dump.pcd", *image = NULL, *device = NULL;
    uint32_t width = 640, height = 480, spp = 100;
    int cam_conf_id = -1;
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "-o")) obj = argv[i + 1];
        else if (!strcmp(argv[i], "-w")) width = (uint32_t)atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "-h")) height = (uint32_t)atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "-s")) spp = (uint32_t)atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "-p")) pcd = argv[i + 1];
        else if (!strcmp(argv[i], "-i")) image = argv[i + 1];
        else if (!strcmp(argv[i], "-c")) cam_conf_id = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "-d")) device = argv[i + 1];
        else { fprintf(stderr, "unknown option: %s\n", argv[i]); return EXIT_FAILURE; }
    }
    if (cam_conf_id < 0) cam_conf_id = image ? 0 : 2;                      /* wrapper.rs:53: cam_conf_id = 2 */

    struct futhark_context_config *cfg = futhark_context_config_new();
    if (device) futhark_context_config_set_device(cfg, device);
    struct futhark_context *ctx = futhark_context_new(cfg);
    if (!ctx) { fprintf(stderr, "lys_save: futhark_context_new failed (no CUDA device?)\n"); return EXIT_FAILURE; }

    size_t n_tris, n_mat_components; float *tri_data; uint32_t *tri_mats; float *mat_data;
    load_obj_data((char *)obj, &n_tris, &n_mat_components, &tri_data, &tri_mats, &mat_data);
    struct futhark_f32_3d *f_tris = futhark_new_f32_3d(ctx, tri_data, (int64_t)n_tris, 3, 3);
    struct futhark_u32_1d *f_tri_mats = futhark_new_u32_1d(ctx, tri_mats, (int64_t)n_tris);
    struct futhark_f32_2d *f_mats = futhark_new_f32_2d(ctx, mat_data, (int64_t)n_mat_components / 28, 28);
    float origin[3] = {0.0f, 0.8f, 1.8f};
    struct futhark_f32_1d *f_origin = futhark_new_f32_1d(ctx, origin, 3);
    if (!f_tris || !f_tri_mats || !f_mats || !f_origin) check(ctx, 1, "futhark_new_*");

    struct futhark_opaque_state *state = NULL;
    check(ctx, futhark_entry_init(ctx, &state, 0, height, width, (uint32_t)cam_conf_id, f_tris, f_tri_mats, f_mats, 0.0f, 0.0f, f_origin),
          "futhark_entry_init");

    size_t n_px = (size_t)width * height;
    if (image) {                                                           /* main.rs:34-49 */
        struct futhark_f32_3d *f_img = NULL;
        check(ctx, futhark_entry_sample_n_frames(ctx, &f_img, state, spp), "futhark_entry_sample_n_frames");
        float *rgb = malloc(n_px * 3 * sizeof(float));
        check(ctx, futhark_values_f32_3d(ctx, f_img, rgb), "futhark_values_f32_3d");
        futhark_free_f32_3d(ctx, f_img);
        if (lys_write_ppm_rgb(image, rgb, width, height)) { perror(image); return EXIT_FAILURE; }
        printf("image %ux%u, %u passes -> %s\n", width, height, spp, image);
        free(rgb);
    } else {                                                               /* main.rs:22-31, wrapper.rs:77-101 */
        struct futhark_opaque_state *new_state = NULL;
        struct futhark_f32_3d *f_points = NULL;
        check(ctx, futhark_entry_sample_points_n(ctx, &new_state, &f_points, state, spp), "futhark_entry_sample_points_n");
        futhark_free_opaque_state(ctx, state);
        state = new_state;
        float *points = malloc(n_px * 4 * sizeof(float));
        check(ctx, futhark_values_f32_3d(ctx, f_points, points), "futhark_values_f32_3d");
        futhark_free_f32_3d(ctx, f_points);
        if (lys_write_pcd_xyz(pcd, points, n_px)) { perror(pcd); return EXIT_FAILURE; }
        printf("points %zu (%ux%u, %u samples per pixel) -> %s\n", n_px, width, height, spp, pcd);
        free(points);
    }

    futhark_free_opaque_state(ctx, state);
    futhark_free_f32_3d(ctx, f_tris); futhark_free_u32_1d(ctx, f_tri_mats); futhark_free_f32_2d(ctx, f_mats); futhark_free_f32_1d(ctx, f_origin);
    futhark_context_free(ctx);
    futhark_context_config_free(cfg);
    free_obj_data(tri_data, tri_mats, mat_data);
    return 0;
}
