// ref_shim.cpp -- compiles the REFERENCE's own CPU implementations, from the sources where they
// lie under /root/reference (nothing is copied into this repo), behind extern "C" wrappers so that
// tests and bench.py --impl reference can call them through ctypes.  Output: oracle/_ref/libsamplenet_ref.so
// (git-ignored, travels to the GPU box).  Test infrastructure, not product.
//
//   registration/src/chamfer_distance/chamfer_distance.cpp : nnsearch (:59-87), chamfer_distance_forward (:90-111),
//                                                             chamfer_distance_backward (:114-177)
//   classification/structural_losses/approxmatch.cpp       : approxmatch_cpu (:17-76), matchcost_cpu (:77-99),
//                                                             matchcostgrad_cpu (:100-125)
#include <cstdio>
#include <cstdlib>

// ---- reference Chamfer (needs the torch stub on the include path) ----
#include "registration/src/chamfer_distance/chamfer_distance.cpp"

// The two CUDA launchers the .cpp declares are never called by the CPU entry points; define them so the
// library links.
int ChamferDistanceKernelLauncher(const int, const int, const float *, const int, const float *, float *, int *, float *, int *) { abort(); }
int ChamferDistanceGradKernelLauncher(const int, const int, const float *, const int, const float *, const float *, const int *, const float *, const int *, float *, float *) { abort(); }

// ---- reference EMD CPU functions: the file has its own main() and GPU launcher declarations ----
#define main approxmatch_reference_main
#define randomf approxmatch_reference_randomf
#include "classification/structural_losses/approxmatch.cpp"
#undef main
#undef randomf
void approxmatchLauncher(int, int, int, const float *, const float *, float *) { abort(); }
void matchcostLauncher(int, int, int, const float *, const float *, const float *, float *) { abort(); }
void matchcostgradLauncher(int, int, int, const float *, const float *, const float *, float *) { abort(); }

static at::Tensor T(const void *p, long a, long b2 = 1, long c = 1) {
    at::Tensor t; t.ptr = const_cast<void *>(p); t.sizes[0] = a; t.sizes[1] = b2; t.sizes[2] = c; t.sizes[3] = 1; return t;
}

extern "C" {
__attribute__((visibility("default")))
void ref_chamfer_forward(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist1, float *dist2, int *idx1, int *idx2) {
    chamfer_distance_forward(T(xyz1, b, n, 3), T(xyz2, b, m, 3), T(dist1, b, n), T(dist2, b, m), T(idx1, b, n), T(idx2, b, m));
}
__attribute__((visibility("default")))
void ref_chamfer_backward(int b, int n, int m, const float *xyz1, const float *xyz2, float *gradxyz1, float *gradxyz2,
                          const float *graddist1, const float *graddist2, const int *idx1, const int *idx2) {
    chamfer_distance_backward(T(xyz1, b, n, 3), T(xyz2, b, m, 3), T(gradxyz1, b, n, 3), T(gradxyz2, b, m, 3),
                              T(graddist1, b, n), T(graddist2, b, m), T(idx1, b, n), T(idx2, b, m));
}
__attribute__((visibility("default")))
void ref_approxmatch_cpu(int b, int n, int m, const float *xyz1, const float *xyz2, float *match) {
    approxmatch_cpu(b, n, m, const_cast<float *>(xyz1), const_cast<float *>(xyz2), match);
}
__attribute__((visibility("default")))
void ref_matchcost_cpu(int b, int n, int m, const float *xyz1, const float *xyz2, const float *match, float *cost) {
    matchcost_cpu(b, n, m, const_cast<float *>(xyz1), const_cast<float *>(xyz2), const_cast<float *>(match), cost);
}
__attribute__((visibility("default")))
void ref_matchcostgrad_cpu(int b, int n, int m, const float *xyz1, const float *xyz2, const float *match, float *grad2) {
    matchcostgrad_cpu(b, n, m, const_cast<float *>(xyz1), const_cast<float *>(xyz2), const_cast<float *>(match), grad2);
}
}
