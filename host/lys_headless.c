/* lys_headless.c -- the reference's interactive host loop without the window.
 *
 * Same call sequence as demo-interactive/liblys.c (reference): load_obj_data (liblys.c:292-295), futhark_new_f32_3d /
 * u32_1d / f32_2d (:297-302), futhark_entry_init with seed 0, camera (0,0.8,1.8) (:133-144), then per frame
 * futhark_entry_step + futhark_entry_render + futhark_values_i32_2d + futhark_free_i32_2d (:107-115), key events through
 * futhark_entry_key (:92-97), FUT_CHECK error handling (liblys.h:32-40).  SDL2 is replaced by a PPM writer, because the
 * reference's libSDL2.a blob is missing and there is no display on the GPU box.  Links against libtracer.a (static, as
 * the reference Makefile:48-49 does) and libljus.
 *
 *   lys_headless -o scene.obj [-w 800] [-h 600] [-n frames] [-k keycodes,comma,separated] [-p out.ppm] [-d device]
 */
#define _POSIX_C_SOURCE 200809L
#include "tracer.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <inttypes.h>

void load_obj_data(char *obj_path, size_t *num_tris, size_t *num_mat_components, float **tri_data, uint32_t **tri_mats, float **mat_data);
void free_obj_data(float *tri_data, uint32_t *tri_mats, float *mat_data);

#define FUT_CHECK(ctx, x) fut_check(ctx, x, __FILE__, __LINE__)
static void fut_check(struct futhark_context *ctx, int res, const char *file, int line) {
    if (res != 0) {
        char *msg = futhark_context_get_error(ctx);
        fprintf(stderr, "%s:%d: Futhark error %d: %s\n", file, line, res, msg ? msg : "");
        free(msg);
        exit(EXIT_FAILURE);
    }
}

int main(int argc, char **argv) {
    const char *obj = NULL, *ppm = NULL, *device = NULL, *keys = "109";   /* SDLK_m: accumulate (lib.fut:154-155) */
    uint32_t width = 800, height = 600;                                   /* INITIAL_WIDTH / INITIAL_HEIGHT liblys.c:18-19 */
    int frames = 4;
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "-o")) obj = argv[i + 1];
        else if (!strcmp(argv[i], "-w")) width = (uint32_t)atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "-h")) height = (uint32_t)atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "-n")) frames = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "-k")) keys = argv[i + 1];
        else if (!strcmp(argv[i], "-p")) ppm = argv[i + 1];
        else if (!strcmp(argv[i], "-d")) device = argv[i + 1];
        else { fprintf(stderr, "unknown option: %s\n", argv[i]); return EXIT_FAILURE; }
    }
    if (!obj) { fprintf(stderr, "usage: %s -o scene.obj [-w W] [-h H] [-n frames] [-k keys] [-p out.ppm] [-d dev]\n", argv[0]); return EXIT_FAILURE; }

    struct futhark_context_config *cfg = futhark_context_config_new();
    if (device) futhark_context_config_set_device(cfg, device);
    struct futhark_context *ctx = futhark_context_new(cfg);
    if (!ctx) { fprintf(stderr, "futhark_context_new failed (no CUDA device?)\n"); return EXIT_FAILURE; }

    size_t num_tris, num_mat_components; float *tri_data; uint32_t *tri_mats; float *mat_data;
    load_obj_data((char *)obj, &num_tris, &num_mat_components, &tri_data, &tri_mats, &mat_data);
    struct futhark_f32_3d *f_tris = futhark_new_f32_3d(ctx, tri_data, (int64_t)num_tris, 3, 3);
    struct futhark_u32_1d *f_mats_ix = futhark_new_u32_1d(ctx, tri_mats, (int64_t)num_tris);
    struct futhark_f32_2d *f_mats = futhark_new_f32_2d(ctx, mat_data, (int64_t)num_mat_components / 28, 28);
    float cam_origin_[3] = {0.0f, 0.8f, 1.8f};
    struct futhark_f32_1d *cam_origin = futhark_new_f32_1d(ctx, cam_origin_, 3);

    struct futhark_opaque_state *state, *next;
    FUT_CHECK(ctx, futhark_entry_init(ctx, &state, 0, height, width, 0, f_tris, f_mats_ix, f_mats, 0.0f, 0.0f, cam_origin));
    /* window_size_updated (liblys.c:38-41) */
    FUT_CHECK(ctx, futhark_entry_resize(ctx, &next, height, width, state));
    futhark_free_opaque_state(ctx, state); state = next;
    /* key-down events (liblys.c:92-97) */
    char *kcopy = strdup(keys);
    for (char *tok = strtok(kcopy, ","); tok; tok = strtok(NULL, ",")) {
        FUT_CHECK(ctx, futhark_entry_key(ctx, &next, 0, (int32_t)strtol(tok, NULL, 0), state));
        futhark_free_opaque_state(ctx, state); state = next;
    }
    free(kcopy);

    int32_t *data = malloc((size_t)width * height * sizeof(int32_t));
    for (int f = 0; f < frames; f++) {                                   /* sdl_loop liblys.c:104-123 */
        struct futhark_i32_2d *out_arr;
        FUT_CHECK(ctx, futhark_entry_step(ctx, &next, state));
        futhark_free_opaque_state(ctx, state); state = next;
        FUT_CHECK(ctx, futhark_entry_render(ctx, &out_arr, state));
        FUT_CHECK(ctx, futhark_values_i32_2d(ctx, out_arr, data));
        FUT_CHECK(ctx, futhark_free_i32_2d(ctx, out_arr));
    }
    if (ppm) {
        FILE *fp = fopen(ppm, "wb");
        if (!fp) { perror(ppm); return EXIT_FAILURE; }
        fprintf(fp, "P6\n%" PRIu32 " %" PRIu32 "\n255\n", width, height);
        for (size_t i = 0; i < (size_t)width * height; i++) {
            uint32_t p = (uint32_t)data[i];                               /* masks 0xFF0000 / 0xFF00 / 0xFF, liblys.c:59 */
            unsigned char rgb[3] = {(unsigned char)(p >> 16), (unsigned char)(p >> 8), (unsigned char)p};
            fwrite(rgb, 1, 3, fp);
        }
        fclose(fp);
    }
    uint64_t sum = 0;
    for (size_t i = 0; i < (size_t)width * height; i++) sum += (uint32_t)data[i] & 0xFFFFFFu;
    printf("frames %d  %ux%u  argb checksum %" PRIu64 "\n", frames, width, height, sum);

    free(data);
    FUT_CHECK(ctx, futhark_free_opaque_state(ctx, state));
    futhark_free_f32_3d(ctx, f_tris); futhark_free_u32_1d(ctx, f_mats_ix); futhark_free_f32_2d(ctx, f_mats); futhark_free_f32_1d(ctx, cam_origin);
    futhark_context_free(ctx);
    futhark_context_config_free(cfg);
    free_obj_data(tri_data, tri_mats, mat_data);
    return 0;
}
