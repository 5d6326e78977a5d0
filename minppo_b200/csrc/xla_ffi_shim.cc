// minppo_b200 -- XLA FFI (jax.ffi) shim over the C ABI in include/minppo_b200.h.
//
// STATUS: NOT BUILT AND NOT TESTED IN THIS IMAGE.  The XLA FFI headers (xla/ffi/api/{c_api,api,ffi}.h)
// ship inside jaxlib and jaxlib is not installed here; build.sh does not compile this
// file.  It is the file a maintainer compiles where JAX exists:
//
//   g++ -O2 -fPIC -shared -std=c++17 -I$(python -c 'import jax; print(jax.ffi.include_dir())') \
//       -I/usr/local/cuda/include -Iinclude minppo_b200/csrc/xla_ffi_shim.cc \
//       -Lminppo_b200/lib -lminppo_b200 -Wl,-rpath,'$ORIGIN' -o minppo_b200/lib/libminppo_b200_xla.so
//
// It contains argument unpacking only: every handler forwards raw device pointers and the XLA
// stream to one extern "C" entry point.  The handlers replace, inside the one jitted program of
// /root/reference/minppo/train.py:
//   MinppoGae     -> _calculate_gae                      (train.py:185-207)
//   MinppoUpdate  -> GAE + the epoch / minibatch scans   (train.py:185-281)
// XLA owns every buffer; in-place updates of params / mu / nu / count are expressed on the Python
// side with input_output_aliases (minppo_b200/jax_ffi.py), so the result pointers below equal the
// aliased operand pointers.
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <string>
#include <utility>

#include "minppo_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error ToError(int code) {
  if (code == MINPPO_OK) return ffi::Error::Success();
  return ffi::Error(code == MINPPO_ERR_ARG ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    std::string("minppo_b200: ") + minppo_last_error());
}

// ---- GAE ----------------------------------------------------------------------------------
static ffi::Error GaeImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> reward, ffi::Buffer<ffi::F32> value,
                          ffi::Buffer<ffi::PRED> done, ffi::Buffer<ffi::F32> last_val, float gamma, float gae_lambda,
                          ffi::ResultBuffer<ffi::F32> adv, ffi::ResultBuffer<ffi::F32> tgt) {
  const auto dims = reward.dimensions();                       // [T, N], time-major (train.py:179)
  if (dims.size() != 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "minppo_gae: reward must be [T, N]");
  return ToError(minppo_gae(reward.typed_data(), value.typed_data(),
                            reinterpret_cast<const uint8_t*>(done.untyped_data()), last_val.typed_data(),
                            adv->typed_data(), tgt->typed_data(), static_cast<int32_t>(dims[0]), dims[1], gamma,
                            gae_lambda, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MinppoGae, GaeImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()   // reward
                                  .Arg<ffi::Buffer<ffi::F32>>()   // value
                                  .Arg<ffi::Buffer<ffi::PRED>>()  // done
                                  .Arg<ffi::Buffer<ffi::F32>>()   // last_val
                                  .Attr<float>("gamma")
                                  .Attr<float>("gae_lambda")
                                  .Ret<ffi::Buffer<ffi::F32>>()   // advantages
                                  .Ret<ffi::Buffer<ffi::F32>>()); // targets

// ---- learner update -------------------------------------------------------------------------
// One context per (device, COMPLETE minppo_config); created on first use, kept for the process.  minppo_ctx_create
// snapshots the whole config (learning rates, gamma, clip_eps, coefficients, anneal_lr, total_timesteps, use_tanh, ...),
// so the key is every byte of it: a second call with the same shapes but different hyper-parameters gets its own
// context instead of silently training with the first one's settings.  Callers value-initialise the struct
// (`minppo_config c = {}`), so padding bytes compare equal.
namespace {
using CtxKey = std::pair<int, std::string>;
std::mutex g_mu;
std::map<CtxKey, minppo_ctx*> g_ctx;

minppo_ctx* GetCtx(const minppo_config& c, int device, int* err) {
  std::lock_guard<std::mutex> lock(g_mu);
  CtxKey key{device, std::string(reinterpret_cast<const char*>(&c), sizeof(c))};
  auto it = g_ctx.find(key);
  if (it != g_ctx.end()) return it->second;
  minppo_ctx* ctx = nullptr;
  *err = minppo_ctx_create(&c, nullptr, &ctx);
  if (*err == MINPPO_OK) g_ctx[key] = ctx;
  return ctx;
}
}  // namespace

static ffi::Error UpdateImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> mu,
                             ffi::Buffer<ffi::F32> nu, ffi::Buffer<ffi::S32> count, ffi::Buffer<ffi::F32> obs,
                             ffi::Buffer<ffi::F32> action, ffi::Buffer<ffi::F32> value, ffi::Buffer<ffi::F32> reward,
                             ffi::Buffer<ffi::F32> log_prob, ffi::Buffer<ffi::PRED> done,
                             ffi::Buffer<ffi::F32> last_val, ffi::Buffer<ffi::U32> rng,
                             // rl.* / training.* / opt.* / model.* (config.py:50-84), passed as static attributes
                             int32_t num_minibatches, int32_t update_epochs, int64_t total_timesteps,
                             bool anneal_lr, int32_t hidden_size, int32_t num_layers, bool use_tanh,
                             int32_t prng_mode, float training_lr, float opt_lr, float max_grad_norm, float gamma,
                             float gae_lambda, float clip_eps, float ent_coef, float vf_coef,
                             ffi::ResultBuffer<ffi::F32> params_out, ffi::ResultBuffer<ffi::F32> mu_out,
                             ffi::ResultBuffer<ffi::F32> nu_out, ffi::ResultBuffer<ffi::S32> count_out,
                             ffi::ResultBuffer<ffi::U32> rng_out, ffi::ResultBuffer<ffi::F32> losses) {
  const auto od = obs.dimensions();                             // [T, N, D]
  const auto ad = action.dimensions();                          // [T, N, A]
  if (od.size() != 3 || ad.size() != 3) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "obs/action must be [T, N, .]");
  // in-place contract: the Python side aliases operands 0..3 to results 0..3
  if (params_out->untyped_data() != params.untyped_data() || mu_out->untyped_data() != mu.untyped_data() ||
      nu_out->untyped_data() != nu.untyped_data() || count_out->untyped_data() != count.untyped_data())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "minppo_update: params/mu/nu/count must be aliased to the results");
  minppo_config c = {};
  c.num_steps = static_cast<int32_t>(od[0]); c.num_envs = static_cast<int32_t>(od[1]);
  c.obs_dim = static_cast<int32_t>(od[2]); c.act_dim = static_cast<int32_t>(ad[2]);
  c.num_minibatches = num_minibatches; c.update_epochs = update_epochs; c.total_timesteps = total_timesteps;
  c.anneal_lr = anneal_lr; c.hidden_size = hidden_size; c.num_layers = num_layers; c.use_tanh = use_tanh;
  c.prng_mode = prng_mode; c.world_size = 1; c.rank = 0; c.fast_tanh = 1;
  c.training_lr = training_lr; c.opt_lr = opt_lr; c.max_grad_norm = max_grad_norm; c.gamma = gamma;
  c.gae_lambda = gae_lambda; c.clip_eps = clip_eps; c.ent_coef = ent_coef; c.vf_coef = vf_coef;
  c.adam_b1 = 0.9; c.adam_b2 = 0.999; c.adam_eps = 1e-5; c.adam_eps_root = 0.0;    // optax.adam(eps=1e-5), train.py:118
  int device = 0, err = 0;
  cudaGetDevice(&device);
  minppo_ctx* ctx = GetCtx(c, device, &err);
  if (!ctx) return ToError(err);
  // use_graph = 0: XLA may itself be capturing this stream into a command buffer
  return ToError(minppo_update(ctx, params_out->typed_data(), mu_out->typed_data(), nu_out->typed_data(),
                               count_out->typed_data(), obs.typed_data(), action.typed_data(), value.typed_data(),
                               reward.typed_data(), log_prob.typed_data(),
                               reinterpret_cast<const uint8_t*>(done.untyped_data()), last_val.typed_data(),
                               rng.typed_data(), rng_out->typed_data(), losses->typed_data(), 0, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    MinppoUpdate, UpdateImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F32>>()   // params arena [P]
        .Arg<ffi::Buffer<ffi::F32>>()   // mu
        .Arg<ffi::Buffer<ffi::F32>>()   // nu
        .Arg<ffi::Buffer<ffi::S32>>()   // count [1]
        .Arg<ffi::Buffer<ffi::F32>>()   // obs
        .Arg<ffi::Buffer<ffi::F32>>()   // action
        .Arg<ffi::Buffer<ffi::F32>>()   // value
        .Arg<ffi::Buffer<ffi::F32>>()   // reward
        .Arg<ffi::Buffer<ffi::F32>>()   // log_prob
        .Arg<ffi::Buffer<ffi::PRED>>()  // done
        .Arg<ffi::Buffer<ffi::F32>>()   // last_val
        .Arg<ffi::Buffer<ffi::U32>>()   // rng [2]
        .Attr<int32_t>("num_minibatches").Attr<int32_t>("update_epochs").Attr<int64_t>("total_timesteps")
        .Attr<bool>("anneal_lr").Attr<int32_t>("hidden_size").Attr<int32_t>("num_layers").Attr<bool>("use_tanh")
        .Attr<int32_t>("prng_mode").Attr<float>("training_lr").Attr<float>("opt_lr").Attr<float>("max_grad_norm")
        .Attr<float>("gamma").Attr<float>("gae_lambda").Attr<float>("clip_eps").Attr<float>("ent_coef")
        .Attr<float>("vf_coef")
        .Ret<ffi::Buffer<ffi::F32>>()   // params'
        .Ret<ffi::Buffer<ffi::F32>>()   // mu'
        .Ret<ffi::Buffer<ffi::F32>>()   // nu'
        .Ret<ffi::Buffer<ffi::S32>>()   // count'
        .Ret<ffi::Buffer<ffi::U32>>()   // rng'
        .Ret<ffi::Buffer<ffi::F32>>()); // losses [E, M, 4]

// ---- rollout policy step ---------------------------------------------------------------------
// Replaces `pi, value = network.apply(params, last_obs); rng, action_rng = split(rng); action = pi.sample(...);
// log_prob = pi.log_prob(action)` (train.py:157-160).  Takes the SAME static attributes as MinppoUpdate so that both
// calls resolve to one context (one set of weight images, one hyper-parameter block).
static ffi::Error PolicyStepImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> last_obs,
                                 ffi::Buffer<ffi::U32> rng, int32_t num_steps, int32_t act_dim,
                                 int32_t num_minibatches, int32_t update_epochs, int64_t total_timesteps,
                                 bool anneal_lr, int32_t hidden_size, int32_t num_layers, bool use_tanh,
                                 int32_t prng_mode, float training_lr, float opt_lr, float max_grad_norm, float gamma,
                                 float gae_lambda, float clip_eps, float ent_coef, float vf_coef, bool weights_current,
                                 ffi::ResultBuffer<ffi::F32> action, ffi::ResultBuffer<ffi::F32> log_prob,
                                 ffi::ResultBuffer<ffi::F32> value, ffi::ResultBuffer<ffi::U32> rng_out) {
  const auto od = last_obs.dimensions();                        // [N, D]
  if (od.size() != 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "last_obs must be [N, D]");
  minppo_config c = {};
  c.num_steps = num_steps; c.num_envs = static_cast<int32_t>(od[0]);
  c.obs_dim = static_cast<int32_t>(od[1]); c.act_dim = act_dim;
  c.num_minibatches = num_minibatches; c.update_epochs = update_epochs; c.total_timesteps = total_timesteps;
  c.anneal_lr = anneal_lr; c.hidden_size = hidden_size; c.num_layers = num_layers; c.use_tanh = use_tanh;
  c.prng_mode = prng_mode; c.world_size = 1; c.rank = 0; c.fast_tanh = 1;
  c.training_lr = training_lr; c.opt_lr = opt_lr; c.max_grad_norm = max_grad_norm; c.gamma = gamma;
  c.gae_lambda = gae_lambda; c.clip_eps = clip_eps; c.ent_coef = ent_coef; c.vf_coef = vf_coef;
  c.adam_b1 = 0.9; c.adam_b2 = 0.999; c.adam_eps = 1e-5; c.adam_eps_root = 0.0;
  int device = 0, err = 0;
  cudaGetDevice(&device);
  minppo_ctx* ctx = GetCtx(c, device, &err);
  if (!ctx) return ToError(err);
  return ToError(minppo_policy_step(ctx, params.typed_data(), last_obs.typed_data(), rng.typed_data(),
                                    rng_out->typed_data(), action->typed_data(), log_prob->typed_data(),
                                    value->typed_data(), nullptr,
                                    weights_current ? MINPPO_POLICY_WEIGHTS_CURRENT : 0, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    MinppoPolicyStep, PolicyStepImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F32>>()   // params arena [P]
        .Arg<ffi::Buffer<ffi::F32>>()   // last_obs [N, D]
        .Arg<ffi::Buffer<ffi::U32>>()   // rng [2]
        .Attr<int32_t>("num_steps").Attr<int32_t>("act_dim")
        .Attr<int32_t>("num_minibatches").Attr<int32_t>("update_epochs").Attr<int64_t>("total_timesteps")
        .Attr<bool>("anneal_lr").Attr<int32_t>("hidden_size").Attr<int32_t>("num_layers").Attr<bool>("use_tanh")
        .Attr<int32_t>("prng_mode").Attr<float>("training_lr").Attr<float>("opt_lr").Attr<float>("max_grad_norm")
        .Attr<float>("gamma").Attr<float>("gae_lambda").Attr<float>("clip_eps").Attr<float>("ent_coef")
        .Attr<float>("vf_coef").Attr<bool>("weights_current")
        .Ret<ffi::Buffer<ffi::F32>>()   // action [N, A]
        .Ret<ffi::Buffer<ffi::F32>>()   // log_prob [N]
        .Ret<ffi::Buffer<ffi::F32>>()   // value [N]
        .Ret<ffi::Buffer<ffi::U32>>()); // rng' [2]

// ---- bootstrap value --------------------------------------------------------------------------
// Replaces `_, last_val = network.apply(runner_state.train_state.params, runner_state.last_obs)` (train.py:182-183): the
// critic-only form of minppo_policy_step (action == log_prob == mean == NULL, no key).  Same static attributes as the
// other two calls, so that all three resolve to one context.
static ffi::Error BootstrapValueImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> last_obs,
                                     int32_t num_steps, int32_t act_dim, int32_t num_minibatches, int32_t update_epochs,
                                     int64_t total_timesteps, bool anneal_lr, int32_t hidden_size, int32_t num_layers,
                                     bool use_tanh, int32_t prng_mode, float training_lr, float opt_lr,
                                     float max_grad_norm, float gamma, float gae_lambda, float clip_eps, float ent_coef,
                                     float vf_coef, bool weights_current, ffi::ResultBuffer<ffi::F32> value) {
  const auto od = last_obs.dimensions();                        // [N, D]
  if (od.size() != 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "last_obs must be [N, D]");
  minppo_config c = {};
  c.num_steps = num_steps; c.num_envs = static_cast<int32_t>(od[0]);
  c.obs_dim = static_cast<int32_t>(od[1]); c.act_dim = act_dim;
  c.num_minibatches = num_minibatches; c.update_epochs = update_epochs; c.total_timesteps = total_timesteps;
  c.anneal_lr = anneal_lr; c.hidden_size = hidden_size; c.num_layers = num_layers; c.use_tanh = use_tanh;
  c.prng_mode = prng_mode; c.world_size = 1; c.rank = 0; c.fast_tanh = 1;
  c.training_lr = training_lr; c.opt_lr = opt_lr; c.max_grad_norm = max_grad_norm; c.gamma = gamma;
  c.gae_lambda = gae_lambda; c.clip_eps = clip_eps; c.ent_coef = ent_coef; c.vf_coef = vf_coef;
  c.adam_b1 = 0.9; c.adam_b2 = 0.999; c.adam_eps = 1e-5; c.adam_eps_root = 0.0;
  int device = 0, err = 0;
  cudaGetDevice(&device);
  minppo_ctx* ctx = GetCtx(c, device, &err);
  if (!ctx) return ToError(err);
  return ToError(minppo_policy_step(ctx, params.typed_data(), last_obs.typed_data(), nullptr, nullptr, nullptr, nullptr,
                                    value->typed_data(), nullptr,
                                    weights_current ? MINPPO_POLICY_WEIGHTS_CURRENT : 0, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    MinppoBootstrapValue, BootstrapValueImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F32>>()   // params arena [P]
        .Arg<ffi::Buffer<ffi::F32>>()   // last_obs [N, D]
        .Attr<int32_t>("num_steps").Attr<int32_t>("act_dim")
        .Attr<int32_t>("num_minibatches").Attr<int32_t>("update_epochs").Attr<int64_t>("total_timesteps")
        .Attr<bool>("anneal_lr").Attr<int32_t>("hidden_size").Attr<int32_t>("num_layers").Attr<bool>("use_tanh")
        .Attr<int32_t>("prng_mode").Attr<float>("training_lr").Attr<float>("opt_lr").Attr<float>("max_grad_norm")
        .Attr<float>("gamma").Attr<float>("gae_lambda").Attr<float>("clip_eps").Attr<float>("ent_coef")
        .Attr<float>("vf_coef").Attr<bool>("weights_current")
        .Ret<ffi::Buffer<ffi::F32>>()); // value [N]
