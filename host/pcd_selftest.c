/* pcd_selftest.c -- runs the writers of lys_pcd.h on raw f32 data from a file (CPU only; tests/test_c_host.py).
 *   pcd_selftest pcd  in.f32 n out.pcd        in.f32 = n x 4 floats (x, y, z, intensity)
 *   pcd_selftest ppm  in.f32 w h out.ppm      in.f32 = h x w x 3 floats
 *   pcd_selftest fmt  in.f32 n                prints one formatted float per line */
#define _POSIX_C_SOURCE 200809L
#include "lys_pcd.h"

static float *slurp(const char *path, size_t count) {
    FILE *fp = fopen(path, "rb");
    if (!fp) { perror(path); exit(EXIT_FAILURE); }
    float *d = malloc((count ? count : 1) * sizeof(float));
    if (fread(d, sizeof(float), count, fp) != count) { fprintf(stderr, "%s: short read\n", path); exit(EXIT_FAILURE); }
    fclose(fp);
    return d;
}

int main(int argc, char **argv) {
    if (argc == 5 && !strcmp(argv[1], "pcd")) {
        size_t n = (size_t)atoll(argv[3]);
        float *d = slurp(argv[2], 4 * n);
        return lys_write_pcd_xyz(argv[4], d, n);
    }
    if (argc == 6 && !strcmp(argv[1], "ppm")) {
        uint32_t w = (uint32_t)atoi(argv[3]), h = (uint32_t)atoi(argv[4]);
        float *d = slurp(argv[2], (size_t)w * h * 3);
        return lys_write_ppm_rgb(argv[5], d, w, h);
    }
    if (argc == 4 && !strcmp(argv[1], "fmt")) {
        size_t n = (size_t)atoll(argv[3]);
        float *d = slurp(argv[2], n);
        char buf[64];
        for (size_t i = 0; i < n; i++) { lys_format_f32(d[i], buf); puts(buf); }
        return 0;
    }
    fprintf(stderr, "usage: pcd_selftest pcd|ppm|fmt ...\n");
    return EXIT_FAILURE;
}
