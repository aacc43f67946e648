// bxg_jax_ffi.cc -- XLA FFI custom-call shim over the C ABI in include/bxg.h.
//
// NOT BUILT in this repository's environment: it needs xla/ffi/api/ffi.h, which
// ships inside jaxlib (API names follow the jax >= 0.4.31 FFI docs
// and are unverified here).  Build where jaxlib is installed:
//
//   g++ -O2 -fPIC -shared -std=c++17 bxg_jax_ffi.cc \
//       -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       -I../../../include -L../.. -lbxg -o libbxg_jax_ffi.so
//
// The handler receives the vmapped (batched) operands as whole buffers
// (vmap_method="broadcast_all" on the Python side), so one
// custom call = one bxg_step launch for the whole env batch.  XLA owns every
// buffer; outputs are pre-allocated by XLA; the stream is XLA's.
#include <cstdint>

#include "bxg.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

// The model handle is created once per (System, device) on the Python side with
// bxg_model_create and passed as an int64 attribute (pointer value); this keeps
// the handler free of global mutable state (pmap enters it from one host thread
// per device).
inline const BxgModel* model_of(int64_t h) { return reinterpret_cast<const BxgModel*>(h); }

template <typename Buf>
float* f32(Buf& b) { return reinterpret_cast<float*>(b.untyped_data()); }

#define BXG_STATE_ARGS(X) \
  X(q) X(qd) X(x_pos) X(x_rot) X(xd_ang) X(xd_vel) X(root_com) X(cinr_pos) X(cinr_rot) X(cinr_i) X(cinr_mass) \
  X(cd_ang) X(cd_vel) X(cdof_ang) X(cdof_vel) X(cdofd_ang) X(cdofd_vel) X(mass_mx) X(mass_mx_inv) X(con_jac) \
  X(con_diag) X(con_aref) X(qf_smooth) X(qf_constraint) X(qdd)

#define BXG_IN(n) ffi::Buffer<ffi::F32> in_##n,
#define BXG_OUT(n) ffi::ResultBuffer<ffi::F32> out_##n,

ffi::Error StepImpl(cudaStream_t stream, int64_t model, int32_t n_frames, BXG_STATE_ARGS(BXG_IN)
                    ffi::Buffer<ffi::F32> act, BXG_STATE_ARGS(BXG_OUT) ffi::ResultBuffer<ffi::S32> status) {
  BxgState in, out;
#define BXG_SET(n) in.n = f32(in_##n); out.n = f32(*out_##n);
  BXG_STATE_ARGS(BXG_SET)
#undef BXG_SET
  const int64_t n_env = in_q.dimensions()[0];
  int rc = bxg_step(model_of(model), n_env, n_frames, &in, f32(act), &out, BXG_STEP_DEFAULT, nullptr, stream);
  if (rc != BXG_OK) return ffi::Error(ffi::ErrorCode::kInternal, bxg_last_error());
  return ffi::Error::Success();
}

ffi::Error InitImpl(cudaStream_t stream, int64_t model, ffi::Buffer<ffi::F32> q, ffi::Buffer<ffi::F32> qd,
                    BXG_STATE_ARGS(BXG_OUT) ffi::ResultBuffer<ffi::S32> status) {
  BxgState out;
#define BXG_SET(n) out.n = f32(*out_##n);
  BXG_STATE_ARGS(BXG_SET)
#undef BXG_SET
  int rc = bxg_init(model_of(model), q.dimensions()[0], f32(q), f32(qd), &out, stream);
  if (rc != BXG_OK) return ffi::Error(ffi::ErrorCode::kInternal, bxg_last_error());
  return ffi::Error::Success();
}

}  // namespace

#define BXG_ARG(n) .Arg<ffi::Buffer<ffi::F32>>()
#define BXG_RET(n) .Ret<ffi::Buffer<ffi::F32>>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(BxgStep, StepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("model")
                                  .Attr<int32_t>("n_frames") BXG_STATE_ARGS(BXG_ARG)
                                  .Arg<ffi::Buffer<ffi::F32>>() BXG_STATE_ARGS(BXG_RET)
                                  .Ret<ffi::Buffer<ffi::S32>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(BxgInit, InitImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("model")
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>() BXG_STATE_ARGS(BXG_RET)
                                  .Ret<ffi::Buffer<ffi::S32>>());
