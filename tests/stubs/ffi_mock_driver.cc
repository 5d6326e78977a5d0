// TEST DRIVER over the mock FFI header (tests/stubs/xla/ffi/api/ffi.h): includes the shim's translation unit and exposes
// plain C entry points that build mock Buffers from raw device pointers and call the shim's handler BODIES -- so that the
// context cache (keyed on the whole minppo_config), the aliasing checks and the argument forwarding run on a GPU.
#include "../../minppo_b200/csrc/xla_ffi_shim.cc"

#include <cstring>

static char g_msg[512];
static int finish(const ffi::Error& e) {
  std::strncpy(g_msg, e.message().c_str(), sizeof(g_msg) - 1);
  return e.success() ? 0 : static_cast<int>(e.code());
}

extern "C" {
const char* mock_last_message() { return g_msg; }
int mock_ctx_count() { std::lock_guard<std::mutex> lock(g_mu); return static_cast<int>(g_ctx.size()); }

int mock_gae(void* stream, float* reward, float* value, void* done, float* last_val, float* adv, float* tgt, int64_t T, int64_t N,
             float gamma, float lam) {
  return finish(GaeImpl(static_cast<cudaStream_t>(stream), ffi::Buffer<ffi::F32>(reward, {T, N}), ffi::Buffer<ffi::F32>(value, {T, N}),
                        ffi::Buffer<ffi::PRED>(done, {T, N}), ffi::Buffer<ffi::F32>(last_val, {N}), gamma, lam,
                        ffi::ResultBuffer<ffi::F32>(ffi::Buffer<ffi::F32>(adv, {T, N})),
                        ffi::ResultBuffer<ffi::F32>(ffi::Buffer<ffi::F32>(tgt, {T, N}))));
}

int mock_update(void* stream, float* params, float* mu, float* nu, int32_t* count, float* obs, float* action, float* value,
                float* reward, float* log_prob, void* done, float* last_val, uint32_t* rng, uint32_t* rng_out, float* losses,
                int64_t T, int64_t N, int64_t D, int64_t A, int64_t P, int32_t num_minibatches, int32_t update_epochs,
                int32_t hidden_size, int32_t num_layers, int32_t anneal_lr, float lr, float clip_eps, int32_t alias_ok) {
  float* p_out = alias_ok ? params : mu;                 // alias_ok == 0: violate the in-place contract on purpose
  return finish(UpdateImpl(
      static_cast<cudaStream_t>(stream), ffi::Buffer<ffi::F32>(params, {P}), ffi::Buffer<ffi::F32>(mu, {P}), ffi::Buffer<ffi::F32>(nu, {P}),
      ffi::Buffer<ffi::S32>(count, {1}), ffi::Buffer<ffi::F32>(obs, {T, N, D}), ffi::Buffer<ffi::F32>(action, {T, N, A}),
      ffi::Buffer<ffi::F32>(value, {T, N}), ffi::Buffer<ffi::F32>(reward, {T, N}), ffi::Buffer<ffi::F32>(log_prob, {T, N}),
      ffi::Buffer<ffi::PRED>(done, {T, N}), ffi::Buffer<ffi::F32>(last_val, {N}), ffi::Buffer<ffi::U32>(rng, {2}),
      num_minibatches, update_epochs, /*total_timesteps=*/1000000000LL, anneal_lr != 0, hidden_size, num_layers, /*use_tanh=*/true,
      /*prng_mode=*/0, /*training_lr=*/lr, /*opt_lr=*/lr, /*max_grad_norm=*/0.5f, /*gamma=*/0.99f, /*gae_lambda=*/0.95f, clip_eps,
      /*ent_coef=*/0.0f, /*vf_coef=*/0.5f,
      ffi::ResultBuffer<ffi::F32>(ffi::Buffer<ffi::F32>(p_out, {P})), ffi::ResultBuffer<ffi::F32>(ffi::Buffer<ffi::F32>(mu, {P})),
      ffi::ResultBuffer<ffi::F32>(ffi::Buffer<ffi::F32>(nu, {P})), ffi::ResultBuffer<ffi::S32>(ffi::Buffer<ffi::S32>(count, {1})),
      ffi::ResultBuffer<ffi::U32>(ffi::Buffer<ffi::U32>(rng_out, {2})),
      ffi::ResultBuffer<ffi::F32>(ffi::Buffer<ffi::F32>(losses, {update_epochs, num_minibatches, 4}))));
}
}  // extern "C"
