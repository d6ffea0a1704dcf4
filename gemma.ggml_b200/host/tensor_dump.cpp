// tensor_dump.cpp — see tensor_dump.h.  Plain C++ file I/O; nothing here touches the GPU.
#include "tensor_dump.h"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <vector>

static std::string g_dir;

void set_tensor_dump_dir(const char *dir) { g_dir = dir ? dir : ""; }

const std::string &tensor_dump_dir() {
    if (g_dir.empty()) {
        const char *e = getenv("GGB_TENSOR_DUMP_DIR");
        g_dir = e && *e ? e : "./tensor_dump";
    }
    return g_dir;
}

static std::string file_of(const char *name, tensor_dump_mode mode) {
    return tensor_dump_dir() + "/" + name + (mode == tensor_dump_mode::target ? "_target" : "_source");
}

bool dump_ptr_data(const char *name, const void *ptr, size_t size, tensor_dump_mode mode) {
    if (!name || (!ptr && size)) return false;
    FILE *f = fopen(file_of(name, mode).c_str(), "wb"); // overwrite, like the reference
    if (!f) return false;
    const bool ok = size == 0 || fwrite(ptr, size, 1, f) == 1;
    return fclose(f) == 0 && ok;
}

void *load_tensor(const char *name, tensor_dump_mode mode, size_t *size) {
    FILE *f = fopen(file_of(name, mode).c_str(), "rb");
    if (!f) return nullptr;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    void *p = n >= 0 ? malloc((size_t)n + 1) : nullptr;
    if (p && n > 0 && fread(p, (size_t)n, 1, f) != 1) {
        free(p);
        p = nullptr;
    }
    fclose(f);
    if (p && size) *size = (size_t)n;
    return p;
}

bool compare_tensors(const char *name) {
    size_t ns = 0, nt = 0;
    void *s = load_tensor(name, tensor_dump_mode::source, &ns), *t = load_tensor(name, tensor_dump_mode::target, &nt);
    const bool same = s && t && ns == nt && memcmp(s, t, ns) == 0;
    free(s);
    free(t);
    return same;
}
bool compare_tensors(const std::string &name) { return compare_tensors(name.c_str()); }

bool compare_tensors_report(const char *name, tensor_compare_report *out) {
    size_t ns = 0, nt = 0;
    void *s = load_tensor(name, tensor_dump_mode::source, &ns), *t = load_tensor(name, tensor_dump_mode::target, &nt);
    const bool ok = s && t && ns == nt && ns % 4 == 0 && out;
    if (ok) {
        tensor_compare_report r;
        r.n = (long long)(ns / 4);
        const float *a = (const float *)s, *b = (const float *)t;
        const uint32_t *ab = (const uint32_t *)s, *bb = (const uint32_t *)t;
        for (long long i = 0; i < r.n; i++) {
            if (ab[i] != bb[i]) r.n_diff++;
            const double d = fabs((double)a[i] - (double)b[i]), m = fabs((double)b[i]);
            if (d > r.max_abs) r.max_abs = d;
            if (m > r.max_target) r.max_target = m;
            const double rel = d / (m > 1e-30 ? m : 1e-30);
            if (d > 0.0 && rel > r.max_rel) r.max_rel = rel;
        }
        *out = r;
    }
    free(s);
    free(t);
    return ok;
}

std::map<std::string, std::string> get_tensor_dump_list() {
    std::map<std::string, std::string> m;
    std::ifstream in(tensor_dump_dir() + "/dump_tensor_list");
    std::string line;
    while (std::getline(in, line)) {
        const size_t c = line.find(':');
        if (c == std::string::npos) continue;
        m[line.substr(0, c)] = line.substr(c + 1);
    }
    return m;
}
