// C ABI of libxlstm_b200.so (see include/xlstm_b200.h): handle, weight binding, state layout, and the
// orchestration of one env step (embed -> L x [LN, proj_up, conv/qkv/gates, state step, proj_down] -> LN ->
// head -> argmax/inv_tokenize). No torch types; no allocation on the step path.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/xlstm_b200.h"
#include "xl_internal.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define XL_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return fail(XL_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                  __LINE__);                                                                   \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct BlockWeights {
  const void* w[XL_W_PER_BLOCK_COUNT];
};

struct GraphCacheEntry {
  cudaGraphExec_t exec = nullptr;
  // key
  void* state = nullptr;
  const void *states = nullptr, *rtg = nullptr, *rewards = nullptr;
  void *tokens = nullptr, *actions = nullptr, *logits = nullptr, *hidden = nullptr;
  int B = 0, mode = 0;
  unsigned flags = 0;
  int64_t launches = 0;
};

}  // namespace

struct xl_handle {
  xl_config cfg;
  int device = 0;
  int num_sms = 148;
  int DH = 0, NCH = 0, Kpad = 0, head_out = 0, num_actions = 0;
  std::vector<BlockWeights> blocks;
  const void* pw[16];  // policy-level weights, index = id - XL_W_POST_NORM
  // workspace (device)
  char* ws = nullptr;
  size_t ws_bytes = 0;
  float *x = nullptr, *xn = nullptr, *u = nullptr, *qkv = nullptr, *act = nullptr, *gate_part = nullptr,
        *gated = nullptr, *partial = nullptr, *s_emb = nullptr, *states_pad = nullptr, *logits = nullptr,
        *d_states = nullptr, *d_rtg = nullptr, *d_rew = nullptr, *d_actions = nullptr, *xtok = nullptr, *hid = nullptr;
  int32_t* d_tokens = nullptr;
  unsigned int* counters = nullptr;
  __nv_bfloat16 *a_hi = nullptr, *a_lo = nullptr;
  int64_t launches = 0;
  std::vector<GraphCacheEntry> graphs;
  size_t a_cap = 0;                    // elements per bf16 hi/lo plane
  int num_tickets = 0;
  cudaStream_t cap_stream = nullptr;   // graph capture never happens on the caller's (possibly legacy) stream
  int state_impl = 1;                  // 1 = TMA ring, 0 = register-batched loads (xl_set_option "state_impl")
  int gemm_impl = 0;                   // 0 auto, 1 CUDA-core, 2 tcgen05      (xl_set_option "gemm_impl")
  int gemm_splitk = 0;                 // cluster split-K in the tcgen05 Linear (xl_set_option "gemm_splitk")
  bool profiling = false;
  std::vector<cudaEvent_t> prof_state;  // start/stop pairs around state-step launches
  std::vector<cudaEvent_t> prof_step;   // start/stop pairs around policy steps
};

namespace {

struct StateLayout {
  size_t c_off, n_off, m_off, conv_off, layer_bytes;
};

StateLayout state_layout(const xl_handle* h, int B) {
  const xl_config& c = h->cfg;
  StateLayout L;
  size_t off = 0;
  L.c_off = off;
  off = align_up(off + sizeof(float) * (size_t)B * c.num_heads * h->DH * h->DH, 256);
  L.n_off = off;
  off = align_up(off + sizeof(float) * (size_t)B * c.num_heads * h->DH, 256);
  L.m_off = off;
  off = align_up(off + sizeof(float) * (size_t)B * c.num_heads, 256);
  L.conv_off = off;
  off = align_up(off + sizeof(float) * (size_t)B * c.conv_kernel * c.inner_dim, 256);
  L.layer_bytes = off;
  return L;
}

int check_batch(const xl_handle* h, int B) {
  if (!h) return fail(XL_ERR_INVALID_ARG, "null handle");
  if (B <= 0 || B > h->cfg.max_batch)
    return fail(XL_ERR_INVALID_ARG, "B=%d outside [1, max_batch=%d]", B, h->cfg.max_batch);
  return XL_OK;
}

// Linear layer dispatch. impl: 0 auto, 1 CUDA-core, 2 tensor-core.
// presplit: the bf16 hi/lo planes of A already sit in h->a_hi / h->a_lo (written by the producing kernel).
int linear(xl_handle* h, const float* A, const void* W, const float* bias, const float* residual, float* out,
           int M, int N, int K, int impl, cudaStream_t s, bool presplit = false) {
  if (K % 8 != 0) return fail(XL_ERR_UNSUPPORTED, "linear: K=%d must be a multiple of 8", K);
  const bool tc_ok = xl::gemm_tc_supported(M, N, K) && (size_t)M * K <= h->a_cap;
  if (impl == 2 && !tc_ok)
    return fail(XL_ERR_UNSUPPORTED, "linear: tcgen05 path needs K %% 64 == 0 and M*K <= %zu (got M=%d K=%d)",
                h->a_cap, M, K);
  if (impl == 2 || (impl == 0 && tc_ok)) {
    if (!presplit) {
      xl::launch_split_bf16(A, K, h->a_hi, h->a_lo, M, K, s);
      h->launches += 1;
    }
    XL_CUDA(xl::launch_gemm_tc(h->a_hi, h->a_lo, (const __nv_bfloat16*)W, bias, residual, out, M, N, K,
                               h->num_sms, h->gemm_splitk ? 0 : 1, s));
    h->launches += 1;
    return XL_OK;
  }
  xl::launch_gemm_simple(A, (const __nv_bfloat16*)W, bias, residual, out, M, N, K, s);
  h->launches += 1;
  XL_CUDA(cudaGetLastError());
  return XL_OK;
}

// One pass of the block stack over M = B*T rows held in h->x (rows ordered [b][t]).
int run_blocks(xl_handle* h, void* state, int B, int T, unsigned flags, cudaStream_t s) {
  const xl_config& c = h->cfg;
  const int d = c.embedding_dim, inner = c.inner_dim, NH = c.num_heads;
  const int M = B * T;
  const StateLayout L = state_layout(h, B);
  const int impl = (flags & XL_FLAG_SIMPLE_GEMM) ? 1 : h->gemm_impl;
  // when the tensor-core Linear will run, the producers write its bf16 hi/lo operand planes directly
  const bool tc_up = impl != 1 && xl::gemm_tc_supported(M, 2 * inner, d) && (size_t)M * d <= h->a_cap;
  const bool tc_down = impl != 1 && xl::gemm_tc_supported(M, d, inner) && (size_t)M * inner <= h->a_cap;
  for (int i = 0; i < c.num_blocks; ++i) {
    const BlockWeights& w = h->blocks[i];
    char* base = (char*)state + (size_t)i * L.layer_bytes;
    // x_n = LN(x) (gamma = 1 + w)
    xl::launch_ln_rows(h->x, d, tc_up ? nullptr : h->xn, d, (const float*)w.w[XL_W_XLSTM_NORM], nullptr, 1,
                       c.ln_eps, M, d, tc_up ? h->a_hi : nullptr, tc_up ? h->a_lo : nullptr, s);
    h->launches += 1;
    // u = x_n @ W_up^T   [M, 2*inner]
    int rc = linear(h, h->xn, w.w[XL_W_PROJ_UP], nullptr, nullptr, h->u, M, 2 * inner, d, impl, s, tc_up);
    if (rc) return rc;
    // conv + silu + q/k/v + gate partials
    xl::ConvQkvParams cp;
    cp.u = h->u;
    cp.conv_state = (float*)(base + L.conv_off);
    cp.conv_w = (const float*)w.w[XL_W_CONV_W];
    cp.conv_b = (const float*)w.w[XL_W_CONV_B];
    cp.wq = (const float*)w.w[XL_W_Q_PROJ];
    cp.wk = (const float*)w.w[XL_W_K_PROJ];
    cp.wv = (const float*)w.w[XL_W_V_PROJ];
    cp.wi = (const float*)w.w[XL_W_IGATE_W];
    cp.wf = (const float*)w.w[XL_W_FGATE_W];
    cp.qk = h->qkv;
    cp.v = h->qkv + (size_t)2 * M * inner;
    cp.act = h->act;
    cp.gate_part = h->gate_part;
    cp.B = B; cp.T = T; cp.inner = inner; cp.NH = NH; cp.KS = c.conv_kernel; cp.NCH = h->NCH;
    if (!xl::launch_conv_qkv_gates(cp, s))
      return fail(XL_ERR_UNSUPPORTED, "conv/qkv kernel not instantiated for KS=%d T=%d NH=%d", cp.KS, cp.T, cp.NH);
    h->launches += 1;
    // state step (+ GroupNorm + skip + output gate)
    xl::StateStepParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.C = (float*)(base + L.c_off);
    sp.n = (float*)(base + L.n_off);
    sp.m = (float*)(base + L.m_off);
    sp.qk = h->qkv;
    sp.v = h->qkv + (size_t)2 * M * inner;
    sp.gate_part = h->gate_part;
    sp.igate_b = (const float*)w.w[XL_W_IGATE_B];
    sp.fgate_b = (const float*)w.w[XL_W_FGATE_B];
    sp.outnorm_w = (const float*)w.w[XL_W_OUTNORM];
    sp.skip = (const float*)w.w[XL_W_SKIP];
    sp.act = h->act;
    sp.u = h->u;
    sp.out = tc_down ? nullptr : h->gated;
    sp.out_hi = tc_down ? h->a_hi : nullptr;
    sp.out_lo = tc_down ? h->a_lo : nullptr;
    sp.partial = h->partial;
    sp.B = B; sp.T = T; sp.NH = NH; sp.DH = h->DH; sp.inner = inner; sp.NCH = h->NCH;
    sp.ln_eps = c.ln_eps; sp.cell_eps = c.cell_eps;
    sp.impl = h->state_impl; sp.num_layers = c.num_blocks;
    cudaEvent_t pe0 = nullptr, pe1 = nullptr;
    if (h->profiling) {
      XL_CUDA(cudaEventCreate(&pe0));
      XL_CUDA(cudaEventCreate(&pe1));
      XL_CUDA(cudaEventRecord(pe0, s));
    }
    XL_CUDA(xl::launch_state_step(sp, h->num_sms, s));
    h->launches += 2;
    if (h->profiling) {
      XL_CUDA(cudaEventRecord(pe1, s));
      h->prof_state.push_back(pe0);
      h->prof_state.push_back(pe1);
    }
    XL_CUDA(xl::launch_state_finalize(sp, h->num_sms, s));
    // x = x + gated @ W_down^T
    rc = linear(h, h->gated, w.w[XL_W_PROJ_DOWN], nullptr, h->x, h->x, M, d, inner, impl, s, tc_down);
    if (rc) return rc;
  }
  XL_CUDA(cudaGetLastError());
  return XL_OK;
}

// Encoder over x_in [B,T,d] -> x_out [B,T,d] (post_blocks_norm applied).
int run_encoder(xl_handle* h, void* state, const float* x_in, float* x_out, int B, int T, int mode,
                unsigned flags, cudaStream_t s) {
  const xl_config& c = h->cfg;
  const int d = c.embedding_dim;
  const float* post_w = (const float*)h->pw[XL_W_POST_NORM - XL_W_POST_NORM];
  if (mode == XL_MODE_FUSED || T == 1) {
    if (x_in != h->x) {
      XL_CUDA(cudaMemcpyAsync(h->x, x_in, sizeof(float) * (size_t)B * T * d, cudaMemcpyDeviceToDevice, s));
    }
    int rc = run_blocks(h, state, B, T, flags, s);
    if (rc) return rc;
    xl::launch_ln_rows(h->x, d, x_out, d, post_w, nullptr, 1, c.ln_eps, B * T, d, nullptr, nullptr, s);
    h->launches += 1;
  } else if (mode == XL_MODE_PER_TOKEN) {
    // reference order: for token: for block  (decision_xlstm.py:161-165). x_in may alias x_out: token t's
    // input row is consumed (gathered) before its output row is written.
    const float* src = x_in;
    if (x_in == x_out) {
      XL_CUDA(cudaMemcpyAsync(h->xtok, x_in, sizeof(float) * (size_t)B * T * d, cudaMemcpyDeviceToDevice, s));
      src = h->xtok;
    }
    for (int t = 0; t < T; ++t) {
      xl::launch_copy_rows(src + (size_t)t * d, (int64_t)T * d, h->x, d, B, d, s);
      h->launches += 1;
      int rc = run_blocks(h, state, B, 1, flags, s);
      if (rc) return rc;
      xl::launch_ln_rows(h->x, d, x_out + (size_t)t * d, (int64_t)T * d, post_w, nullptr, 1, c.ln_eps, B, d,
                         nullptr, nullptr, s);
      h->launches += 1;
    }
  } else {
    return fail(XL_ERR_INVALID_ARG, "unknown mode %d", mode);
  }
  XL_CUDA(cudaGetLastError());
  return XL_OK;
}

int run_policy(xl_handle* h, void* state, const float* states, const float* rtg, const float* rewards,
               int32_t* tokens, float* actions, float* logits, float* hidden, int B, int mode, unsigned flags,
               cudaStream_t s) {
  const xl_config& c = h->cfg;
  const int d = c.embedding_dim, T = c.tokens_per_step;
  const int impl = (flags & XL_FLAG_SIMPLE_GEMM) ? 1 : h->gemm_impl;
  auto PW = [&](int id) { return h->pw[id - XL_W_POST_NORM]; };
  // embed_state: Linear(204 -> d) on zero-padded K
  xl::launch_pad_rows(states, c.state_dim, h->states_pad, h->Kpad, B, s);
  h->launches += 1;
  int rc = linear(h, h->states_pad, PW(XL_W_EMBED_STATE_W), (const float*)PW(XL_W_EMBED_STATE_B), nullptr,
                  h->s_emb, B, d, h->Kpad, impl, s);
  if (rc) return rc;
  float* xt = (mode == XL_MODE_FUSED) ? h->x : h->xtok;
  xl::launch_embed_tokens(h->s_emb, rtg, rewards, (const float*)PW(XL_W_EMBED_RETURN_W),
                          (const float*)PW(XL_W_EMBED_RETURN_B), (const float*)PW(XL_W_EMBED_REWARD_W),
                          (const float*)PW(XL_W_EMBED_REWARD_B), (const float*)PW(XL_W_EMBED_LN_W),
                          (const float*)PW(XL_W_EMBED_LN_B), c.embed_ln_eps, xt, B, d, s);
  h->launches += 1;
  float* hid = hidden ? hidden : h->hid;
  rc = run_encoder(h, state, xt, hid, B, T, mode, flags, s);
  if (rc) return rc;
  // action head on the rtg token (row b*T + pos): gather to a dense [B,d] then Linear(d -> 2192 / 274)
  float* xa = h->s_emb;  // reuse [B,d]
  xl::launch_copy_rows(hid + (size_t)c.action_token_pos * d, (int64_t)T * d, xa, d, B, d, s);
  h->launches += 1;
  const bool discrete = (flags & XL_FLAG_DISCRETE) != 0;
  // discrete branch only needs the first num_actions logits (multi_domain_discrete_dt_model.py:99-101)
  const int n_out = discrete ? h->num_actions : h->head_out;
  float* lg = logits ? logits : h->logits;
  if (discrete) {
    // keep the row pitch of the full head so both branches share the argmax kernel's addressing
    rc = linear(h, xa, PW(XL_W_HEAD_W), (const float*)PW(XL_W_HEAD_B), nullptr, h->logits, B, n_out, d, impl, s);
    if (rc) return rc;
    if (logits) {
      XL_CUDA(cudaMemcpyAsync(logits, h->logits, sizeof(float) * (size_t)B * n_out, cudaMemcpyDeviceToDevice, s));
    }
    xl::launch_argmax_tokens(h->logits, n_out, B, c.act_dim, n_out, c.discrete_actions, 1, 0.f, 0.f, tokens,
                             actions, s);
  } else {
    rc = linear(h, xa, PW(XL_W_HEAD_W), (const float*)PW(XL_W_HEAD_B), nullptr, lg, B, n_out, d, impl, s);
    if (rc) return rc;
    const float bw = (c.tok_max_val - c.tok_min_val) / (float)c.action_channels;
    xl::launch_argmax_tokens(lg, h->head_out, B, c.act_dim, h->num_actions, c.discrete_actions, 0, bw,
                             c.tok_min_val, tokens, actions, s);
  }
  h->launches += 1;
  XL_CUDA(cudaGetLastError());
  return XL_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int xl_abi_version(void) { return XL_ABI_VERSION; }
const char* xl_last_error(void) { return g_err; }

int xl_create(const xl_config* cfg, xl_handle** out) {
  if (!cfg || !out) return fail(XL_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  const xl_config& c = *cfg;
  if (c.embedding_dim <= 0 || c.embedding_dim % 8) return fail(XL_ERR_UNSUPPORTED, "embedding_dim %% 8 != 0");
  if (c.num_heads <= 0 || c.num_heads > 8) return fail(XL_ERR_UNSUPPORTED, "num_heads must be in [1,8]");
  if (c.inner_dim <= 0 || c.inner_dim % (4 * c.num_heads) || c.inner_dim % 8)
    return fail(XL_ERR_UNSUPPORTED, "inner_dim must be a multiple of 8 and of 4*num_heads");
  if (c.qkv_blocksize != 4) return fail(XL_ERR_UNSUPPORTED, "qkv_blocksize must be 4");
  if (c.conv_kernel < 2 || c.conv_kernel > 4) return fail(XL_ERR_UNSUPPORTED, "conv_kernel must be in [2,4]");
  if (c.tokens_per_step < 1 || c.tokens_per_step > 4) return fail(XL_ERR_UNSUPPORTED, "tokens_per_step in [1,4]");
  if (c.action_token_pos < 0 || c.action_token_pos >= c.tokens_per_step)
    return fail(XL_ERR_INVALID_ARG, "action_token_pos out of range");
  if (c.max_batch <= 0 || c.num_blocks <= 0) return fail(XL_ERR_INVALID_ARG, "max_batch/num_blocks <= 0");
  if (c.tokens_per_step != 3)
    return fail(XL_ERR_UNSUPPORTED, "policy token layout is (s, rtg, r): tokens_per_step must be 3");
  const int DH = c.inner_dim / c.num_heads;
  if (DH > 1024) return fail(XL_ERR_UNSUPPORTED, "head_dim > 1024");

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(XL_ERR_NO_DEVICE, "no CUDA device: the xlstm_b200 path has no CPU fallback");
  }
  xl_handle* h = new (std::nothrow) xl_handle();
  if (!h) return fail(XL_ERR_INVALID_ARG, "out of host memory");
  h->cfg = c;
  h->DH = DH;
  {
    // gate pre-activations are produced as NCH partial sums (64 four-channel blocks per chunk), summed in
    // fixed order by the state kernels
    const int nblk = c.inner_dim / 4;
    int nch = (nblk + 63) / 64;
    if (nch < 1) nch = 1;
    if (nch > 16) nch = 16;
    h->NCH = nch;
    if ((nblk + nch - 1) / nch > 128) { delete h; return fail(XL_ERR_UNSUPPORTED, "inner_dim too large"); }
  }
  h->Kpad = (int)align_up((size_t)c.state_dim, 64);
  h->num_actions = c.discrete_actions + c.action_channels;
  h->head_out = h->num_actions * c.act_dim;
  h->blocks.resize(c.num_blocks);
  for (auto& b : h->blocks) memset(&b, 0, sizeof(b));
  memset(h->pw, 0, sizeof(h->pw));
  cudaGetDevice(&h->device);
  cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device);

  // workspace carve-up
  const size_t M = (size_t)c.max_batch * 4;  // up to 4 tokens per step
  const size_t B = (size_t)c.max_batch;
  const size_t d = c.embedding_dim, inner = c.inner_dim;
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  const size_t o_x = carve(4 * M * d), o_xn = carve(4 * M * d), o_xtok = carve(4 * M * d);
  const size_t o_hid = carve(4 * M * d);
  const size_t o_u = carve(4 * M * 2 * inner), o_qkv = carve(4 * M * 3 * inner), o_act = carve(4 * M * inner);
  const size_t o_gp = carve(4 * M * 16 * 2 * c.num_heads), o_gated = carve(4 * M * inner);
  const size_t o_part = carve(4 * B * c.num_heads * 32 * 4 * DH);  // RS <= 32, T <= 4
  const size_t o_semb = carve(4 * B * d), o_sp = carve(4 * B * h->Kpad), o_lg = carve(4 * B * h->head_out);
  const size_t o_ds = carve(4 * B * c.state_dim), o_dr = carve(4 * B), o_dw = carve(4 * B);
  const size_t o_da = carve(4 * B * c.act_dim), o_dt = carve(4 * B * c.act_dim);
  h->num_tickets = 4096;
  const size_t o_cnt = carve(4 * (size_t)h->num_tickets);

  size_t a_cap = M * (inner > d ? inner : d);
  if (a_cap < B * (size_t)h->Kpad) a_cap = B * (size_t)h->Kpad;
  if (a_cap < (size_t)1 << 20) a_cap = (size_t)1 << 20;   // room for xl_linear unit tests
  h->a_cap = a_cap;
  const size_t o_hi = carve(2 * a_cap), o_lo = carve(2 * a_cap);
  h->ws_bytes = off;
  cudaError_t e = cudaMalloc((void**)&h->ws, h->ws_bytes);
  if (e != cudaSuccess) {
    delete h;
    return fail(XL_ERR_CUDA, "cudaMalloc(workspace %zu B) failed: %s", off, cudaGetErrorString(e));
  }
  cudaMemset(h->ws, 0, h->ws_bytes);
  cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking);
  h->x = (float*)(h->ws + o_x); h->xn = (float*)(h->ws + o_xn); h->xtok = (float*)(h->ws + o_xtok);
  h->hid = (float*)(h->ws + o_hid);
  h->u = (float*)(h->ws + o_u); h->qkv = (float*)(h->ws + o_qkv); h->act = (float*)(h->ws + o_act);
  h->gate_part = (float*)(h->ws + o_gp); h->gated = (float*)(h->ws + o_gated);
  h->partial = (float*)(h->ws + o_part); h->s_emb = (float*)(h->ws + o_semb);
  h->states_pad = (float*)(h->ws + o_sp); h->logits = (float*)(h->ws + o_lg);
  h->d_states = (float*)(h->ws + o_ds); h->d_rtg = (float*)(h->ws + o_dr); h->d_rew = (float*)(h->ws + o_dw);
  h->d_actions = (float*)(h->ws + o_da); h->d_tokens = (int32_t*)(h->ws + o_dt);
  h->counters = (unsigned int*)(h->ws + o_cnt);

  h->a_hi = (__nv_bfloat16*)(h->ws + o_hi); h->a_lo = (__nv_bfloat16*)(h->ws + o_lo);
  *out = h;
  return XL_OK;
}

void xl_destroy(xl_handle* h) {
  if (!h) return;
  for (auto& g : h->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  if (h->ws) cudaFree(h->ws);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  delete h;
}

int xl_state_dim_padded(const xl_handle* h) { return h ? h->Kpad : 0; }

int xl_bind_weight(xl_handle* h, int layer, int which, const void* dev_ptr, int dtype, int64_t numel) {
  if (!h || !dev_ptr) return fail(XL_ERR_INVALID_ARG, "null argument");
  const xl_config& c = h->cfg;
  const int64_t d = c.embedding_dim, inner = c.inner_dim, NH = c.num_heads, KS = c.conv_kernel;
  int64_t want = -1;
  int want_dtype = 0;
  if (layer >= 0) {
    if (layer >= c.num_blocks) return fail(XL_ERR_INVALID_ARG, "layer %d >= num_blocks", layer);
    switch (which) {
      case XL_W_XLSTM_NORM: want = d; break;
      case XL_W_PROJ_UP: want = 2 * inner * d; want_dtype = 1; break;
      case XL_W_Q_PROJ: case XL_W_K_PROJ: case XL_W_V_PROJ: want = inner * 4; break;
      case XL_W_CONV_W: want = inner * KS; break;
      case XL_W_CONV_B: want = inner; break;
      case XL_W_IGATE_W: case XL_W_FGATE_W: want = NH * 3 * inner; break;
      case XL_W_IGATE_B: case XL_W_FGATE_B: want = NH; break;
      case XL_W_OUTNORM: case XL_W_SKIP: want = inner; break;
      case XL_W_PROJ_DOWN: want = d * inner; want_dtype = 1; break;
      default: return fail(XL_ERR_INVALID_ARG, "unknown per-block weight id %d", which);
    }
  } else {
    switch (which) {
      case XL_W_POST_NORM: want = d; break;
      case XL_W_EMBED_STATE_W: want = d * h->Kpad; want_dtype = 1; break;
      case XL_W_EMBED_STATE_B: case XL_W_EMBED_RETURN_W: case XL_W_EMBED_RETURN_B:
      case XL_W_EMBED_REWARD_W: case XL_W_EMBED_REWARD_B: case XL_W_EMBED_LN_W: case XL_W_EMBED_LN_B:
        want = d; break;
      case XL_W_HEAD_W: want = (int64_t)h->head_out * d; want_dtype = 1; break;
      case XL_W_HEAD_B: want = h->head_out; break;
      default: return fail(XL_ERR_INVALID_ARG, "unknown policy weight id %d", which);
    }
  }
  if (numel != want) return fail(XL_ERR_INVALID_ARG, "weight %d (layer %d): numel %lld, expected %lld", which,
                                 layer, (long long)numel, (long long)want);
  if (dtype != want_dtype)
    return fail(XL_ERR_INVALID_ARG, "weight %d (layer %d): dtype %d, expected %d (0=fp32,1=bf16)", which, layer,
                dtype, want_dtype);
  if (((uintptr_t)dev_ptr) % 16) return fail(XL_ERR_INVALID_ARG, "weight pointer must be 16-byte aligned");
  if (layer >= 0) h->blocks[layer].w[which] = dev_ptr;
  else h->pw[which - XL_W_POST_NORM] = dev_ptr;
  // bound pointers are baked into cached graphs
  for (auto& g : h->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  h->graphs.clear();
  return XL_OK;
}

int xl_weights_ready(const xl_handle* h) {
  if (!h) return fail(XL_ERR_INVALID_ARG, "null handle");
  for (int i = 0; i < h->cfg.num_blocks; ++i)
    for (int k = 0; k < XL_W_PER_BLOCK_COUNT; ++k)
      if (!h->blocks[i].w[k]) return fail(XL_ERR_NOT_READY, "block %d weight id %d not bound", i, k);
  for (int id = XL_W_POST_NORM; id <= XL_W_HEAD_B; ++id)
    if (!h->pw[id - XL_W_POST_NORM]) return fail(XL_ERR_NOT_READY, "policy weight id %d not bound", id);
  return XL_OK;
}

size_t xl_state_bytes(const xl_handle* h, int B) {
  if (!h || B <= 0) return 0;
  return state_layout(h, B).layer_bytes * (size_t)h->cfg.num_blocks;
}

int xl_state_layout(const xl_handle* h, int B, int layer, int part, size_t* offset_bytes, size_t* size_bytes) {
  if (!h || !offset_bytes || !size_bytes) return fail(XL_ERR_INVALID_ARG, "null argument");
  if (layer < 0 || layer >= h->cfg.num_blocks || B <= 0) return fail(XL_ERR_INVALID_ARG, "bad layer/B");
  const StateLayout L = state_layout(h, B);
  const xl_config& c = h->cfg;
  size_t o, sz;
  switch (part) {
    case XL_STATE_C: o = L.c_off; sz = sizeof(float) * (size_t)B * c.num_heads * h->DH * h->DH; break;
    case XL_STATE_N: o = L.n_off; sz = sizeof(float) * (size_t)B * c.num_heads * h->DH; break;
    case XL_STATE_M: o = L.m_off; sz = sizeof(float) * (size_t)B * c.num_heads; break;
    case XL_STATE_CONV: o = L.conv_off; sz = sizeof(float) * (size_t)B * c.conv_kernel * c.inner_dim; break;
    default: return fail(XL_ERR_INVALID_ARG, "bad state part %d", part);
  }
  *offset_bytes = (size_t)layer * L.layer_bytes + o;
  *size_bytes = sz;
  return XL_OK;
}

int xl_state_reset(xl_handle* h, void* state, const uint8_t* env_mask, int B, void* stream) {
  int rc = check_batch(h, B);
  if (rc) return rc;
  if (!state) return fail(XL_ERR_INVALID_ARG, "null state");
  cudaStream_t s = (cudaStream_t)stream;
  const StateLayout L = state_layout(h, B);
  const xl_config& c = h->cfg;
  if (!env_mask) {
    XL_CUDA(cudaMemsetAsync(state, 0, L.layer_bytes * (size_t)c.num_blocks, s));
    return XL_OK;
  }
  for (int i = 0; i < c.num_blocks; ++i) {
    char* base = (char*)state + (size_t)i * L.layer_bytes;
    xl::launch_state_reset((float*)(base + L.c_off), (float*)(base + L.n_off), (float*)(base + L.m_off),
                           (float*)(base + L.conv_off), env_mask, B, (int64_t)c.num_heads * h->DH * h->DH,
                           (int64_t)c.num_heads * h->DH, c.num_heads, (int64_t)c.conv_kernel * c.inner_dim, s);
    h->launches += 1;
  }
  XL_CUDA(cudaGetLastError());
  return XL_OK;
}

int xl_encoder_step(xl_handle* h, void* state, const float* x_in, float* x_out, int B, int T, int mode,
                    unsigned flags, void* stream) {
  int rc = check_batch(h, B);
  if (rc) return rc;
  if (!state || !x_in || !x_out) return fail(XL_ERR_INVALID_ARG, "null argument");
  if (T < 1 || T > 4) return fail(XL_ERR_UNSUPPORTED, "T=%d outside [1,4]", T);
  rc = xl_weights_ready(h);
  if (rc) return rc;
  return run_encoder(h, state, x_in, x_out, B, T, mode, flags, (cudaStream_t)stream);
}

int xl_mlstm_cell_step(xl_handle* h, float* C, float* n, float* m, const float* qkv, const float* igate,
                       const float* fgate, const float* outnorm_w, float* h_norm, float* h_raw, int B, int T,
                       int rows_split, int cols_per_cta, void* stream) {
  int rc = check_batch(h, B);
  if (rc) return rc;
  if (!C || !n || !m || !qkv || !igate || !fgate || !outnorm_w || !h_norm)
    return fail(XL_ERR_INVALID_ARG, "null argument");
  if (T < 1 || T > 4) return fail(XL_ERR_UNSUPPORTED, "T=%d outside [1,4]", T);
  if (rows_split > 32) return fail(XL_ERR_UNSUPPORTED, "rows_split > 32");
  const xl_config& c = h->cfg;
  cudaStream_t s = (cudaStream_t)stream;
  // pack the caller's gate pre-activations as a single "chunk": [M, 1, 2*NH]
  const int M = B * T, NH = c.num_heads;
  XL_CUDA(cudaMemcpy2DAsync(h->gate_part, sizeof(float) * 2 * NH, igate, sizeof(float) * NH, sizeof(float) * NH,
                            M, cudaMemcpyDeviceToDevice, s));
  XL_CUDA(cudaMemcpy2DAsync(h->gate_part + NH, sizeof(float) * 2 * NH, fgate, sizeof(float) * NH,
                            sizeof(float) * NH, M, cudaMemcpyDeviceToDevice, s));
  xl::StateStepParams sp;
  memset(&sp, 0, sizeof(sp));
  xl::launch_repack_qkv(qkv, h->qkv, h->qkv + (size_t)2 * M * c.inner_dim, M, c.inner_dim, s);
  sp.C = C; sp.n = n; sp.m = m; sp.qk = h->qkv; sp.v = h->qkv + (size_t)2 * M * c.inner_dim;
  sp.gate_part = h->gate_part;
  sp.outnorm_w = outnorm_w; sp.out = h_norm; sp.h_raw = h_raw;
  sp.partial = h->partial;
  sp.B = B; sp.T = T; sp.NH = NH; sp.DH = h->DH; sp.inner = c.inner_dim; sp.NCH = 1;
  sp.rows_split = rows_split; sp.cols_per_cta = cols_per_cta;
  sp.ln_eps = c.ln_eps; sp.cell_eps = c.cell_eps;
  sp.impl = h->state_impl; sp.num_layers = c.num_blocks;
  XL_CUDA(xl::launch_state_step(sp, h->num_sms, s));
  XL_CUDA(xl::launch_state_finalize(sp, h->num_sms, s));
  h->launches += 3;
  return XL_OK;
}

int xl_policy_step(xl_handle* h, void* state, const float* states, const float* rtg, const float* rewards,
                   int32_t* tokens, float* actions, float* logits, float* hidden, int B, int mode,
                   unsigned flags, void* stream) {
  int rc = check_batch(h, B);
  if (rc) return rc;
  if (!state || !states || !rtg || !tokens || !actions) return fail(XL_ERR_INVALID_ARG, "null argument");
  if (mode != XL_MODE_FUSED && mode != XL_MODE_PER_TOKEN) return fail(XL_ERR_INVALID_ARG, "unknown mode %d", mode);
  cudaStream_t s = (cudaStream_t)stream;
  if (!(flags & XL_FLAG_GRAPH)) {
    rc = xl_weights_ready(h);
    if (rc) return rc;
    cudaEvent_t pe0 = nullptr, pe1 = nullptr;
    if (h->profiling) {
      XL_CUDA(cudaEventCreate(&pe0));
      XL_CUDA(cudaEventCreate(&pe1));
      XL_CUDA(cudaEventRecord(pe0, s));
    }
    rc = run_policy(h, state, states, rtg, rewards, tokens, actions, logits, hidden, B, mode, flags, s);
    if (h->profiling) {
      cudaEventRecord(pe1, s);
      h->prof_step.push_back(pe0);
      h->prof_step.push_back(pe1);
    }
    return rc;
  }
  // ---- CUDA-graph replay: capture once per distinct argument tuple --------------------------------
  for (auto& g : h->graphs) {
    if (g.state == state && g.states == states && g.rtg == rtg && g.rewards == rewards && g.tokens == tokens &&
        g.actions == actions && g.logits == logits && g.hidden == hidden && g.B == B && g.mode == mode &&
        g.flags == flags) {
      XL_CUDA(cudaGraphLaunch(g.exec, s));
      h->launches += g.launches;
      return XL_OK;
    }
  }
  rc = xl_weights_ready(h);
  if (rc) return rc;
  // make sure every lazy attribute (dynamic smem opt-in) is set before capture: one eager warm-up is NOT
  // done here because it would advance the state; attributes are set inside launch paths, which is legal
  // during capture (cudaFuncSetAttribute is not a stream operation).
  GraphCacheEntry g;
  g.state = state; g.states = states; g.rtg = rtg; g.rewards = rewards; g.tokens = tokens; g.actions = actions;
  g.logits = logits; g.hidden = hidden; g.B = B; g.mode = mode; g.flags = flags;
  cudaGraph_t graph = nullptr;
  const int64_t before = h->launches;
  XL_CUDA(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
  rc = run_policy(h, state, states, rtg, rewards, tokens, actions, logits, hidden, B, mode, flags, h->cap_stream);
  cudaError_t e = cudaStreamEndCapture(h->cap_stream, &graph);
  g.launches = h->launches - before;
  h->launches = before;
  if (rc) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) return fail(XL_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
  e = cudaGraphInstantiate(&g.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return fail(XL_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
  if (h->graphs.size() >= 8) {
    cudaGraphExecDestroy(h->graphs.front().exec);
    h->graphs.erase(h->graphs.begin());
  }
  h->graphs.push_back(g);
  XL_CUDA(cudaGraphLaunch(g.exec, s));
  h->launches += g.launches;
  return XL_OK;
}

int xl_policy_step_host(xl_handle* h, void* state, const float* h_states, const float* h_rtg,
                        const float* h_rewards, int32_t* h_tokens, float* h_actions, int B, int mode,
                        unsigned flags, void* stream) {
  int rc = check_batch(h, B);
  if (rc) return rc;
  if (!state || !h_states || !h_rtg || !h_tokens || !h_actions) return fail(XL_ERR_INVALID_ARG, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  const xl_config& c = h->cfg;
  XL_CUDA(cudaMemcpyAsync(h->d_states, h_states, sizeof(float) * (size_t)B * c.state_dim, cudaMemcpyHostToDevice, s));
  XL_CUDA(cudaMemcpyAsync(h->d_rtg, h_rtg, sizeof(float) * (size_t)B, cudaMemcpyHostToDevice, s));
  if (h_rewards) XL_CUDA(cudaMemcpyAsync(h->d_rew, h_rewards, sizeof(float) * (size_t)B, cudaMemcpyHostToDevice, s));
  rc = xl_policy_step(h, state, h->d_states, h->d_rtg, h_rewards ? h->d_rew : nullptr, h->d_tokens, h->d_actions,
                      nullptr, nullptr, B, mode, flags, s);
  if (rc) return rc;
  XL_CUDA(cudaMemcpyAsync(h_tokens, h->d_tokens, sizeof(int32_t) * (size_t)B * c.act_dim, cudaMemcpyDeviceToHost, s));
  XL_CUDA(cudaMemcpyAsync(h_actions, h->d_actions, sizeof(float) * (size_t)B * c.act_dim, cudaMemcpyDeviceToHost, s));
  XL_CUDA(cudaStreamSynchronize(s));
  return XL_OK;
}

int xl_linear(xl_handle* h, const float* A, const void* W_bf16, const float* bias, const float* residual,
              float* out, int M, int N, int K, int impl, void* stream) {
  if (!h || !A || !W_bf16 || !out) return fail(XL_ERR_INVALID_ARG, "null argument");
  if (M <= 0 || N <= 0 || K <= 0) return fail(XL_ERR_INVALID_ARG, "bad GEMM shape");
  return linear(h, A, W_bf16, bias, residual, out, M, N, K, impl, (cudaStream_t)stream);
}

int xl_set_option(xl_handle* h, const char* name, int value) {
  if (!h || !name) return fail(XL_ERR_INVALID_ARG, "null argument");
  if (!strcmp(name, "state_impl")) {
    if (value != 0 && value != 1) return fail(XL_ERR_INVALID_ARG, "state_impl must be 0 or 1");
    h->state_impl = value;
  } else if (!strcmp(name, "gemm_splitk")) {
    h->gemm_splitk = value ? 1 : 0;
  } else if (!strcmp(name, "gemm_impl")) {
    if (value < 0 || value > 2) return fail(XL_ERR_INVALID_ARG, "gemm_impl must be 0, 1 or 2");
    h->gemm_impl = value;
  } else {
    return fail(XL_ERR_INVALID_ARG, "unknown option %s", name);
  }
  for (auto& g : h->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  h->graphs.clear();
  return XL_OK;
}

int xl_profile_begin(xl_handle* h) {
  if (!h) return fail(XL_ERR_INVALID_ARG, "null handle");
  h->profiling = true;
  return XL_OK;
}

int xl_profile_end(xl_handle* h, double* state_kernel_ms, int64_t* state_kernel_launches, double* step_ms) {
  if (!h || !state_kernel_ms || !state_kernel_launches || !step_ms) return fail(XL_ERR_INVALID_ARG, "null argument");
  h->profiling = false;
  double a = 0.0, b = 0.0;
  auto drain = [](std::vector<cudaEvent_t>& v, double& acc) -> cudaError_t {
    cudaError_t err = cudaSuccess;
    for (size_t i = 0; i + 1 < v.size(); i += 2) {
      float ms = 0.f;
      cudaError_t e = cudaEventSynchronize(v[i + 1]);
      if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, v[i], v[i + 1]);
      if (e != cudaSuccess) err = e;
      acc += ms;
      cudaEventDestroy(v[i]);
      cudaEventDestroy(v[i + 1]);
    }
    return err;
  };
  const int64_t n = (int64_t)h->prof_state.size() / 2;
  cudaError_t e1 = drain(h->prof_state, a);
  cudaError_t e2 = drain(h->prof_step, b);
  h->prof_state.clear();
  h->prof_step.clear();
  *state_kernel_ms = a;
  *state_kernel_launches = n;
  *step_ms = b;
  if (e1 != cudaSuccess || e2 != cudaSuccess) return fail(XL_ERR_CUDA, "profile events: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
  return XL_OK;
}

int64_t xl_launch_count(xl_handle* h) {
  if (!h) return 0;
  const int64_t v = h->launches;
  h->launches = 0;
  return v;
}

}  // extern "C"
