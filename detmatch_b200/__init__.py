"""detmatch_b200 -- B200-native point-cloud front end behind the MMDetection3D / DetMatch op API.

Only the hot path named in BASELINE.json is implemented: hard / dynamic voxelization
(``mmdet3d.ops.voxel``) and point-in-rotated-box assignment
(``mmdet3d.ops.roiaware_pool3d.points_in_boxes_*``).  All compute runs in hand-written sm_100a
CUDA kernels reached through the C ABI of ``include/pcfe.h`` (``detmatch_b200/lib/libpcfe.so``);
there is no CPU fallback: importing the ops without the built library raises.
"""
__version__ = "0.1.0"
