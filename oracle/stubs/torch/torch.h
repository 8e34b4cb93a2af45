// Minimal stand-in for <torch/torch.h>, used ONLY to compile the reference's
// registration/src/chamfer_distance/chamfer_distance.cpp unmodified with plain g++ (no libtorch)
// into oracle/_ref/.  It provides just what that file touches: at::Tensor::size(i), ::data<T>()
// and a PYBIND11_MODULE that expands to an unused function.  Test infrastructure, not product.
#pragma once
#include <cstdint>
namespace at {
struct Tensor {
    void *ptr;
    int64_t sizes[4];
    int64_t size(int i) const { return sizes[i]; }
    template <typename T> T *data() const { return static_cast<T *>(ptr); }
};
}  // namespace at
struct orc_fake_module {
    template <typename F> void def(const char *, F, const char *) {}
};
#define TORCH_EXTENSION_NAME cd
#define PYBIND11_MODULE(name, m) static void orc_unused_pybind_##name(orc_fake_module &m)
