// Same enumerators, same numeric order as the reference's include/fast_gicp/gicp/gicp_settings.hpp:6-10 and
// fast_vgicp_cuda.hpp:21, lsq_registration.hpp:13 (they cross the C ABI as ints).
#pragma once
namespace fast_gicp {
enum class RegularizationMethod { NONE, MIN_EIG, NORMALIZED_MIN_EIG, PLANE, FROBENIUS };
enum class NeighborSearchMethod { DIRECT27, DIRECT7, DIRECT1, /* supported on only VGICP_CUDA */ DIRECT_RADIUS };
enum class VoxelAccumulationMode { ADDITIVE, ADDITIVE_WEIGHTED, MULTIPLICATIVE };
enum class NearestNeighborMethod { CPU_PARALLEL_KDTREE, GPU_BRUTEFORCE, GPU_RBF_KERNEL };
enum class LSQ_OPTIMIZER_TYPE { GaussNewton, LevenbergMarquardt };
}  // namespace fast_gicp
