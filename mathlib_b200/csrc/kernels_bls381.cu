// BLS381 instantiation of the batch kernels (see kernels.cuh / msm.cuh / launch.cuh).
#define B200_INSTANTIATE 1
#include "launch.cuh"
namespace b200 {
const CurveVTable* vtable_bls381() { return Launch<BLS381>::table(); }
}  // namespace b200
