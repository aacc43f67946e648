// bxg_inst.cu -- one kernel variant per translation unit (-DBXG_VARIANT=k) so
// the variants compile in parallel.  Exposes the two entry points of variant k
// to bxg_api.cu as plain function pointers.
// kernel ids 10 and 11: variants 0 and 1 specialised for the packed layout of Ant / Humanoid (constexpr Dims)
#if BXG_VARIANT == 10
#define BXG_CONST_DIMS "gen/bxg_dims_ant.h"
#define BXG_CONST_DIMS_FN const_dims_ant
#elif BXG_VARIANT == 11
#define BXG_CONST_DIMS "gen/bxg_dims_humanoid.h"
#define BXG_CONST_DIMS_FN const_dims_humanoid
#endif
#include "bxg_kernels.cuh"

#ifndef BXG_VARIANT
#error "compile with -DBXG_VARIANT=0..11"
#endif

namespace {
#if BXG_VARIANT == 0 || BXG_VARIANT == 10
using Cfg = bxg::KernelCfg<16, 4, 6>;
#elif BXG_VARIANT == 1 || BXG_VARIANT == 11
using Cfg = bxg::KernelCfg<32, 6, 7>;
#elif BXG_VARIANT == 2
using Cfg = bxg::KernelCfg<32, 8, 8>;
#elif BXG_VARIANT == 4
using Cfg = bxg::KernelCfg<16, 6, 7>;
#elif BXG_VARIANT == 5
using Cfg = bxg::KernelCfg<32, 4, 16>;
#elif BXG_VARIANT == 6
using Cfg = bxg::KernelCfg<32, 6, 20>;
#elif BXG_VARIANT == 7
using Cfg = bxg::KernelCfg<4, 1, 1>;
#elif BXG_VARIANT == 8
using Cfg = bxg::KernelCfg<4, 2, 2>;
#elif BXG_VARIANT == 9
using Cfg = bxg::KernelCfg<4, 2, 4>;
#else
using Cfg = bxg::KernelCfg<32, 0, 0>;
#endif
}  // namespace

#define BXG_CAT2(a, b) a##b
#define BXG_CAT(a, b) BXG_CAT2(a, b)
// two step kernels per variant: Newton-Schulz (the reference's matrix_inv) and exact Cholesky inverse
// and each of them with lean state I/O (BXG_STEP_LEAN)
extern "C" const void* BXG_CAT(bxg_step_kernel_v, BXG_VARIANT)() { return (const void*)bxg::step_kernel<Cfg, 0, false, BXG_VARIANT>; }
extern "C" const void* BXG_CAT(bxg_step_chol_kernel_v, BXG_VARIANT)() { return (const void*)bxg::step_kernel<Cfg, 1, false, BXG_VARIANT>; }
extern "C" const void* BXG_CAT(bxg_step_lean_kernel_v, BXG_VARIANT)() { return (const void*)bxg::step_kernel<Cfg, 0, true, BXG_VARIANT>; }
extern "C" const void* BXG_CAT(bxg_step_chol_lean_kernel_v, BXG_VARIANT)() { return (const void*)bxg::step_kernel<Cfg, 1, true, BXG_VARIANT>; }
extern "C" const void* BXG_CAT(bxg_init_kernel_v, BXG_VARIANT)() { return (const void*)bxg::init_kernel<Cfg, BXG_VARIANT>; }
// per-env models (bxg_model_create_batched): Newton-Schulz only, on the Ant-class, Humanoid-class and generic variants
#if BXG_VARIANT == 0 || BXG_VARIANT == 1 || BXG_VARIANT == 3
extern "C" const void* BXG_CAT(bxg_step_perenv_kernel_v, BXG_VARIANT)() { return (const void*)bxg::step_kernel<Cfg, 0, false, BXG_VARIANT, true>; }
extern "C" const void* BXG_CAT(bxg_step_perenv_lean_kernel_v, BXG_VARIANT)() { return (const void*)bxg::step_kernel<Cfg, 0, true, BXG_VARIANT, true>; }
extern "C" const void* BXG_CAT(bxg_init_perenv_kernel_v, BXG_VARIANT)() { return (const void*)bxg::init_kernel<Cfg, BXG_VARIANT, true>; }
#endif
