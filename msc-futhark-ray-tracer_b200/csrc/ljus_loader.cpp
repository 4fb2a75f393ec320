/* ljus_loader.cpp -- OBJ/MTL -> (tri_data, tri_mats, mat_data), the host-side scene loader.
 *
 * Replaces the Rust crate `ljus` (reference ljus/src/lib.rs:12-105, built on tobj 0.1.12,
 * ljus/Cargo.toml:8) behind the same C symbols the unchanged C host binds
 * (demo-interactive/liblys.h:14-18):
 *     void load_obj_data(char *obj_path, size_t *num_tris, size_t *num_mat_components,
 *                        float **tri_data, uint32_t **tri_mats, float **mat_data);
 *     void free_obj_data(float *tri_data, uint32_t *tri_mats, float *mat_data);
 *
 * Semantics restated (tobj 0.1.12 is not vendored in the reference tree):
 *  - faces are emitted in file order; a triangle is kept, a quad a b c d becomes (a,b,c),(a,c,d),
 *    an n-gon becomes the fan (v0,v1,v2),(v0,v2,v3),...; negative indices are relative to the
 *    vertices read so far; only the position index of `v/vt/vn` is used;
 *  - every face carries the material active at that point (`usemtl`), lib.rs:45-52; a face with no
 *    active material is an error ("Mesh doesn't have material", lib.rs:45);
 *  - materials keep `newmtl` order; a 28-float row per material (lib.rs:55-102):
 *    [0..12) colour knots = `Sp` padded with (-1,0) pairs, else (610,Kd.r, 550,Kd.g, 460,Kd.b, -1,0 x3);
 *    [12] `Pr` (1), [13] `Pm` (0), [14] `Ni` (tobj default 1.0), [15] `Tf` (1),
 *    [16..28) emission knots = `Em`, else the same RGB layout from `Ke` (0,0,0).
 *  - the line "no of triangles: N" is printed like lib.rs:103.
 * Errors abort the process with a message (the Rust code panics).
 */
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace {

struct Mtl {
    float kd[3] = {0, 0, 0};
    float ni = 1.0f;
    std::map<std::string, std::string> unknown;
};

[[noreturn]] void die(const std::string &msg) {
    fprintf(stderr, "ljus: %s\n", msg.c_str());
    abort();
}

std::vector<std::string> split_ws(const std::string &s) {
    std::vector<std::string> out; std::istringstream is(s); std::string w;
    while (is >> w) out.push_back(w);
    return out;
}
float parse_f32(const std::string &s) {
    char *end = nullptr;
    float v = strtof(s.c_str(), &end);          /* correctly rounded, like Rust's str::parse::<f32> */
    if (end == s.c_str() || *end != '\0') die("bad float '" + s + "'");
    return v;
}
std::vector<float> parse_vec(const std::string &s) {
    std::vector<float> v; for (auto &w : split_ws(s)) v.push_back(parse_f32(w)); return v;
}
std::string dirname_of(const std::string &p) {
    size_t k = p.find_last_of('/');
    return k == std::string::npos ? std::string() : p.substr(0, k + 1);
}

void load_mtl(const std::string &path, std::vector<Mtl> &mats, std::map<std::string, int> &mat_map) {
    std::ifstream f(path);
    if (!f) die("cannot open mtl '" + path + "'");
    std::string line; int cur = -1;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        auto w = split_ws(line);
        if (w.empty() || w[0] == "#") continue;
        const std::string &key = w[0];
        if (key == "newmtl") {
            if (w.size() < 2) die("newmtl without a name");
            mats.emplace_back(); cur = (int)mats.size() - 1; mat_map[w[1]] = cur;
            continue;
        }
        if (cur < 0) continue;
        Mtl &m = mats[cur];
        if (key == "Kd") { if (w.size() < 4) die("Kd needs 3 values"); for (int k = 0; k < 3; k++) m.kd[k] = parse_f32(w[1 + k]); }
        else if (key == "Ni") { if (w.size() < 2) die("Ni needs a value"); m.ni = parse_f32(w[1]); }
        else if (key == "Ka" || key == "Ks" || key == "Ns" || key == "d" || key == "illum" ||
                 key == "map_Ka" || key == "map_Kd" || key == "map_Ks" || key == "map_Ns" || key == "map_d") { /* known to tobj, unused by ljus */ }
        else {
            size_t pos = line.find(key);
            std::string rest = line.substr(pos + key.size());
            size_t a = rest.find_first_not_of(" \t"), b = rest.find_last_not_of(" \t");
            m.unknown[key] = (a == std::string::npos) ? std::string() : rest.substr(a, b - a + 1);
        }
    }
}

void spectrum_row(const Mtl &m, const char *knots_key, const float rgb[3], float *out12) {
    auto it = m.unknown.find(knots_key);
    if (it != m.unknown.end()) {                 /* get_spectrum lib.rs:134-144 */
        std::vector<float> v = parse_vec(it->second);
        for (int i = 0; i < 12; i++) out12[i] = (i < (int)v.size()) ? v[i] : ((i - (int)v.size()) % 2 == 0 ? -1.0f : 0.0f);
        return;
    }
    const float row[12] = {610.0f, rgb[0], 550.0f, rgb[1], 460.0f, rgb[2], -1, 0, -1, 0, -1, 0};
    memcpy(out12, row, sizeof(row));
}
float scalar_or(const Mtl &m, const char *key, float dflt) {
    auto it = m.unknown.find(key);
    return it == m.unknown.end() ? dflt : parse_f32(it->second);
}

int resolve_index(const std::string &tok, int n_vertices) {
    std::string first = tok.substr(0, tok.find('/'));
    char *end = nullptr; long v = strtol(first.c_str(), &end, 10);
    if (end == first.c_str()) die("bad face index '" + tok + "'");
    long ix = v < 0 ? (long)n_vertices + v : v - 1;
    if (ix < 0 || ix >= n_vertices) die("face index out of range '" + tok + "'");
    return (int)ix;
}

void load(const std::string &obj_path, std::vector<float> &tris, std::vector<uint32_t> &tri_mats, std::vector<float> &mat_rows) {
    std::ifstream f(obj_path);
    if (!f) die("Load obj file: cannot open '" + obj_path + "'");
    std::vector<float> pos; std::vector<Mtl> mats; std::map<std::string, int> mat_map;
    int cur_mat = -1; std::string line;
    auto emit = [&](int a, int b, int c) {
        if (cur_mat < 0) die("Mesh doesn't have material");
        tri_mats.push_back((uint32_t)cur_mat);
        for (int v : {a, b, c}) for (int k = 0; k < 3; k++) tris.push_back(pos[3 * (size_t)v + k]);
    };
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        auto w = split_ws(line);
        if (w.empty() || w[0][0] == '#') continue;
        if (w[0] == "v") { if (w.size() < 4) die("v needs 3 values"); for (int k = 0; k < 3; k++) pos.push_back(parse_f32(w[1 + k])); }
        else if (w[0] == "f") {
            if (w.size() < 4) die("face with fewer than 3 vertices");
            int nv = (int)(pos.size() / 3);
            std::vector<int> ix; for (size_t k = 1; k < w.size(); k++) ix.push_back(resolve_index(w[k], nv));
            for (size_t k = 1; k + 1 < ix.size(); k++) emit(ix[0], ix[k], ix[k + 1]);
        }
        else if (w[0] == "mtllib") { if (w.size() < 2) die("mtllib without a file"); load_mtl(dirname_of(obj_path) + w[1], mats, mat_map); }
        else if (w[0] == "usemtl") {
            if (w.size() < 2) die("usemtl without a name");
            auto it = mat_map.find(w[1]); cur_mat = (it == mat_map.end()) ? -1 : it->second;
        }
    }
    mat_rows.assign(mats.size() * 28, 0.0f);
    for (size_t i = 0; i < mats.size(); i++) {
        float *row = &mat_rows[28 * i]; const Mtl &m = mats[i];
        spectrum_row(m, "Sp", m.kd, row);
        row[12] = scalar_or(m, "Pr", 1.0f); row[13] = scalar_or(m, "Pm", 0.0f); row[14] = m.ni; row[15] = scalar_or(m, "Tf", 1.0f);
        float ke[3] = {0, 0, 0};
        auto it = m.unknown.find("Ke");
        if (it != m.unknown.end()) { auto v = parse_vec(it->second); if (v.size() != 3) die("Expected 3-vector parameter"); for (int k = 0; k < 3; k++) ke[k] = v[k]; }
        spectrum_row(m, "Em", ke, row + 16);
    }
    printf("no of triangles: %zu\n", tris.size() / 9);
}

template <class T> T *to_heap(const std::vector<T> &v) {
    T *p = (T *)malloc(sizeof(T) * (v.empty() ? 1 : v.size()));
    if (!p) die("out of memory");
    if (!v.empty()) memcpy(p, v.data(), sizeof(T) * v.size());
    return p;
}

} // namespace

extern "C" {

void load_obj_data(char *obj_path, size_t *num_tris, size_t *num_mat_components, float **tri_data,
                   uint32_t **tri_mats, float **mat_data) {
    std::vector<float> tris, mats; std::vector<uint32_t> tm;
    load(obj_path, tris, tm, mats);
    *num_tris = tm.size(); *num_mat_components = mats.size();
    *tri_data = to_heap(tris); *tri_mats = to_heap(tm); *mat_data = to_heap(mats);
}

void free_obj_data(float *tri_data, uint32_t *tri_mats, float *mat_data) {
    free(tri_data); free(tri_mats); free(mat_data);
}

} // extern "C"
