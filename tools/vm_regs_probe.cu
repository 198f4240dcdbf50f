// Development aid: compiles ONLY vm_pairing_kernel<BLS381, 2, WARPS> so that register / spill counts of a VM change are
// visible in seconds:   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xptxas -v -DPROBE_WARPS=12 -c tools/vm_regs_probe.cu
#include "../mathlib_b200/csrc/pairing_vm.cuh"
#ifndef PROBE_WARPS
#define PROBE_WARPS 4
#endif
#ifndef CURVE
#define CURVE BLS381
#endif
namespace b200 {
template __global__ void vm_pairing_kernel<CURVE, 2, PROBE_WARPS>(size_t, const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*,
                                                             uint8_t*, uint32_t, int*, const uint32_t*, const VmDirEntry*);
}
