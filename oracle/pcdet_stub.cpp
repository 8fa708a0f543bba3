// oracle/pcdet_stub.cpp -- TEST INFRASTRUCTURE ONLY (ours).  Link-time stand-ins for the CUDA
// launchers that thirdparty/Spconv-OpenPCDet/pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:19-27
// declares and roiaware_pool3d_kernel.cu defines: oracle/build_ref.py compiles only the .cpp (for its
// points_in_boxes_cpu), so the GPU entry points of that module must never be reached.
#include <cstdio>
#include <cstdlib>

static void no_cuda(const char* what) {
  std::fprintf(stderr, "pcdet_ref_cpu: %s needs the reference's CUDA kernels, which are not built\n", what);
  std::abort();
}

void roiaware_pool3d_launcher(int, int, int, int, int, int, int, const float*, const float*, const float*, int*,
                              int*, float*, int) {
  no_cuda("roiaware_pool3d_launcher");
}
void roiaware_pool3d_backward_launcher(int, int, int, int, int, int, const int*, const int*, const float*, float*,
                                       int) {
  no_cuda("roiaware_pool3d_backward_launcher");
}
void points_in_boxes_launcher(int, int, int, const float*, const float*, int*) { no_cuda("points_in_boxes_launcher"); }
