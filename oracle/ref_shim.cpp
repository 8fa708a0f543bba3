// oracle/ref_shim.cpp -- pybind shim around the reference's UNMODIFIED CPU sources.
// TEST INFRASTRUCTURE ONLY.  This file is ours; the two reference translation units
// (mmdet3d/ops/voxel/src/voxelization_cpu.cpp, mmdet3d/ops/roiaware_pool3d/src/
// points_in_boxes_cpu.cpp) are compiled from where they lie under /root/reference by
// oracle/build_ref.py and are never copied into this repository.
//
// The declarations below restate the signatures at voxelization.h:8-18 and
// roiaware_pool3d.cpp:40-41 so the linker can resolve them.
#include <torch/extension.h>

#include <vector>

namespace voxelization {
int hard_voxelize_cpu(const at::Tensor &points, at::Tensor &voxels, at::Tensor &coors,
                      at::Tensor &num_points_per_voxel, const std::vector<float> voxel_size,
                      const std::vector<float> coors_range, const int max_points,
                      const int max_voxels, const int NDim);
void dynamic_voxelize_cpu(const at::Tensor &points, at::Tensor &coors,
                          const std::vector<float> voxel_size,
                          const std::vector<float> coors_range, const int NDim);
}  // namespace voxelization

int points_in_boxes_cpu(at::Tensor boxes_tensor, at::Tensor pts_tensor,
                        at::Tensor pts_indices_tensor);

static int hard_voxelize(const at::Tensor &points, at::Tensor &voxels, at::Tensor &coors,
                         at::Tensor &num_points_per_voxel, const std::vector<float> voxel_size,
                         const std::vector<float> coors_range, const int max_points,
                         const int max_voxels, const int NDim) {
  return voxelization::hard_voxelize_cpu(points, voxels, coors, num_points_per_voxel, voxel_size,
                                         coors_range, max_points, max_voxels, NDim);
}

static void dynamic_voxelize(const at::Tensor &points, at::Tensor &coors,
                             const std::vector<float> voxel_size,
                             const std::vector<float> coors_range, const int NDim) {
  voxelization::dynamic_voxelize_cpu(points, coors, voxel_size, coors_range, NDim);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("hard_voxelize", &hard_voxelize, "reference hard_voxelize_cpu", py::arg("points"),
        py::arg("voxels"), py::arg("coors"), py::arg("num_points_per_voxel"),
        py::arg("voxel_size"), py::arg("coors_range"), py::arg("max_points"),
        py::arg("max_voxels"), py::arg("NDim") = 3);
  m.def("dynamic_voxelize", &dynamic_voxelize, "reference dynamic_voxelize_cpu",
        py::arg("points"), py::arg("coors"), py::arg("voxel_size"), py::arg("coors_range"),
        py::arg("NDim") = 3);
  m.def("points_in_boxes_cpu", &points_in_boxes_cpu, "reference points_in_boxes_cpu",
        py::arg("boxes"), py::arg("points"), py::arg("out"));
}
