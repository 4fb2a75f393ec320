/* lys_pcd.h -- the two file formats of the reference's demo-save host, in C.
 *
 * (1) ASCII .pcd point cloud with x y z fields, one record per pixel, as demo-save/src/main.rs:23-31 writes it through
 *     pcd-rs 0.6 (`WriterBuilder::new(points.len(), 1, Default::default(), DataKind::ASCII)` on a `Vec3 {x, y, z}` record,
 *     wrapper.rs:12-18).  pcd-rs is not vendored under the reference; the header below is the PCL v0.7 layout every PCD
 *     reader (PCL, pcd-rs) accepts.  Numbers are printed the way Rust's `Display for f32` prints them: the shortest
 *     decimal string that parses back to the same f32, positional notation, `inf` / `-inf` / `NaN`.
 * (2) 8-bit RGB image of a `sample_n_frames` result: `(x.clamp(0, 1) * 255.99) as u8` per component (the commented-out
 *     capture path, main.rs:34-49).  The reference hands the bytes to the `image` crate (PNG); here they go into a
 *     binary PPM, which needs no library.
 */
#ifndef LYS_PCD_H
#define LYS_PCD_H

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* Shortest round-trip decimal of an f32 in positional notation (Rust `{}`); buf must hold 64 bytes. */
static inline void lys_format_f32(float v, char *buf) {
    if (isnan(v)) { strcpy(buf, "NaN"); return; }
    if (isinf(v)) { strcpy(buf, v < 0 ? "-inf" : "inf"); return; }
    if (v == 0.0f) { strcpy(buf, signbit(v) ? "-0" : "0"); return; }
    char sci[32];
    int prec;
    for (prec = 0; prec < 9; prec++) {                       /* prec digits after the first: 1..9 significant digits */
        snprintf(sci, sizeof sci, "%.*e", prec, (double)v);
        if (strtof(sci, NULL) == v) break;
    }
    if (prec == 9) snprintf(sci, sizeof sci, "%.8e", (double)v);
    /* sci = [-]d[.ddd]e[+-]XX  ->  digits + decimal exponent */
    char digits[16]; int nd = 0, neg = 0;
    const char *p = sci;
    if (*p == '-') { neg = 1; p++; }
    for (; *p && *p != 'e'; p++) if (*p != '.') digits[nd++] = *p;
    int e10 = atoi(p + 1);
    while (nd > 1 && digits[nd - 1] == '0') nd--;           /* 1.50e+00 never happens at the shortest precision, but be safe */
    char *o = buf;
    if (neg) *o++ = '-';
    if (e10 >= nd - 1) {                                     /* integer: digits then zeros */
        memcpy(o, digits, (size_t)nd); o += nd;
        for (int k = 0; k < e10 - (nd - 1); k++) *o++ = '0';
    } else if (e10 >= 0) {                                   /* point inside the digits */
        memcpy(o, digits, (size_t)e10 + 1); o += e10 + 1;
        *o++ = '.';
        memcpy(o, digits + e10 + 1, (size_t)(nd - e10 - 1)); o += nd - e10 - 1;
    } else {                                                 /* 0.000ddd */
        *o++ = '0'; *o++ = '.';
        for (int k = 0; k < -e10 - 1; k++) *o++ = '0';
        memcpy(o, digits, (size_t)nd); o += nd;
    }
    *o = '\0';
}

/* points: [n][4] = (x, y, z, intensity) as futhark_entry_sample_points_n returns them (lib.fut:35-63); only the
 * position is written (wrapper.rs:96-101).  Returns 0 on success. */
static inline int lys_write_pcd_xyz(const char *path, const float *points, size_t n) {
    FILE *fp = fopen(path, "wb");
    if (!fp) return 1;
    fprintf(fp, "# .PCD v.7 - Point Cloud Data file format\nVERSION .7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n");
    fprintf(fp, "WIDTH %zu\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %zu\nDATA ascii\n", n, n);
    char x[64], y[64], z[64];
    for (size_t i = 0; i < n; i++) {
        lys_format_f32(points[4 * i + 0], x); lys_format_f32(points[4 * i + 1], y); lys_format_f32(points[4 * i + 2], z);
        fprintf(fp, "%s %s %s\n", x, y, z);
    }
    return fclose(fp) != 0;
}

/* rgb: [h][w][3] f32 (futhark_entry_sample_n_frames) -> P6 PPM with the reference's quantisation (main.rs:43-46). */
static inline unsigned char lys_quantise_255_99(float x) {
    float c = x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x);      /* f32::clamp; NaN stays NaN and `as u8` saturates it to 0 */
    if (isnan(c)) return 0;
    return (unsigned char)(c * 255.99f);
}
static inline int lys_write_ppm_rgb(const char *path, const float *rgb, uint32_t w, uint32_t h) {
    FILE *fp = fopen(path, "wb");
    if (!fp) return 1;
    fprintf(fp, "P6\n%u %u\n255\n", w, h);
    for (size_t i = 0; i < (size_t)w * h * 3; i++) fputc(lys_quantise_255_99(rgb[i]), fp);
    return fclose(fp) != 0;
}

#endif /* LYS_PCD_H */
