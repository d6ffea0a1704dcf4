// tensor_dump.h — the reference's parity harness (ref: src/tensor_dump.h:26-32, src/tensor_dump.cpp:11-117), rewired:
// same entry points and file naming ("<dir>/<name>_source" | "<name>_target", raw tensor bytes), but
//   * the directory is a run-time setting (the reference hard-codes /home/geraltigas/..., src/tensor_dump.h:12),
//   * tensors are handed over as (pointer, byte size) - dump_ptr_data's signature - because the B200 path has no
//     ggml_tensor (device intermediates are fetched with gemma_model::get_capture first),
//   * besides the reference's byte compare (compare_tensors -> bool) there is the report BASELINE.json asks for:
//     max-abs / max-rel error and the number of differing elements of an fp32 tensor pair.
#ifndef GGB_TENSOR_DUMP_H
#define GGB_TENSOR_DUMP_H

#include <cstddef>
#include <map>
#include <string>

enum class tensor_dump_mode { target = 1, source = 0 };

struct tensor_compare_report {
    long long n = 0, n_diff = 0; // fp32 elements compared, elements whose bits differ
    double max_abs = 0.0;        // max |source - target|
    double max_rel = 0.0;        // max |source - target| / max(|target|, tiny)
    double max_target = 0.0;     // max |target| (the scale tolerances are quoted against)
};

void set_tensor_dump_dir(const char *dir); // default: $GGB_TENSOR_DUMP_DIR, else ./tensor_dump
const std::string &tensor_dump_dir();
// write `size` bytes as "<dir>/<name>_source" or "_target" (ref: dump_tensor / dump_ptr_data); false on I/O error
bool dump_ptr_data(const char *name, const void *ptr, size_t size, tensor_dump_mode mode = tensor_dump_mode::source);
// byte-for-byte equality of "<name>_source" and "<name>_target" (ref: compare_tensors, src/tensor_dump.cpp:33-89)
bool compare_tensors(const char *name);
bool compare_tensors(const std::string &name);
// the same pair read as fp32: error statistics; false if a file is missing or the sizes differ
bool compare_tensors_report(const char *name, tensor_compare_report *out);
// malloc'ed copy of a dumped tensor (ref: load_tensor); caller frees; *size receives the byte count
void *load_tensor(const char *name, tensor_dump_mode mode, size_t *size = nullptr);
// "name:alias" lines of <dir>/dump_tensor_list (ref: get_tensor_dump_list, src/tensor_dump.cpp:95-117)
std::map<std::string, std::string> get_tensor_dump_list();

#endif
