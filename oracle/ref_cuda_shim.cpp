// oracle/ref_cuda_shim.cpp -- pybind shim around the reference's UNMODIFIED CUDA sources
// (mmdet3d/ops/voxel/src/voxelization_cuda.cu, mmdet3d/ops/roiaware_pool3d/src/points_in_boxes_cuda.cu),
// compiled for sm_100a from where they lie under /root/reference by oracle/build_ref_cuda.py.
// TEST / MEASUREMENT INFRASTRUCTURE ONLY: the GPU-vs-GPU column of bench.py ("reference_cuda") and
// tests/native/bench_ops.py.  This file is ours; the declarations restate voxelization.h:24-34 and
// roiaware_pool3d.cpp:32-38 so that the linker can resolve them.
#include <torch/extension.h>

#include <vector>

namespace voxelization {
int hard_voxelize_gpu(const at::Tensor &points, at::Tensor &voxels, at::Tensor &coors,
                      at::Tensor &num_points_per_voxel, const std::vector<float> voxel_size,
                      const std::vector<float> coors_range, const int max_points,
                      const int max_voxels, const int NDim);
void dynamic_voxelize_gpu(const at::Tensor &points, at::Tensor &coors,
                          const std::vector<float> voxel_size,
                          const std::vector<float> coors_range, const int NDim);
}  // namespace voxelization

int points_in_boxes_gpu(at::Tensor boxes_tensor, at::Tensor pts_tensor, at::Tensor box_idx_of_points_tensor);
int points_in_boxes_batch(at::Tensor boxes_tensor, at::Tensor pts_tensor, at::Tensor box_idx_of_points_tensor);

static int hard_voxelize(const at::Tensor &points, at::Tensor &voxels, at::Tensor &coors,
                         at::Tensor &num_points_per_voxel, const std::vector<float> voxel_size,
                         const std::vector<float> coors_range, const int max_points,
                         const int max_voxels, const int NDim) {
  return voxelization::hard_voxelize_gpu(points, voxels, coors, num_points_per_voxel, voxel_size,
                                         coors_range, max_points, max_voxels, NDim);
}

static void dynamic_voxelize(const at::Tensor &points, at::Tensor &coors,
                             const std::vector<float> voxel_size,
                             const std::vector<float> coors_range, const int NDim) {
  voxelization::dynamic_voxelize_gpu(points, coors, voxel_size, coors_range, NDim);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("hard_voxelize", &hard_voxelize, "reference hard_voxelize_gpu", py::arg("points"),
        py::arg("voxels"), py::arg("coors"), py::arg("num_points_per_voxel"),
        py::arg("voxel_size"), py::arg("coors_range"), py::arg("max_points"),
        py::arg("max_voxels"), py::arg("NDim") = 3);
  m.def("dynamic_voxelize", &dynamic_voxelize, "reference dynamic_voxelize_gpu",
        py::arg("points"), py::arg("coors"), py::arg("voxel_size"), py::arg("coors_range"),
        py::arg("NDim") = 3);
  m.def("points_in_boxes_gpu", &points_in_boxes_gpu, "reference points_in_boxes_gpu");
  m.def("points_in_boxes_batch", &points_in_boxes_batch, "reference points_in_boxes_batch");
}
