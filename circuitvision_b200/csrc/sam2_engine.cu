// SAM 2.1 image path (Hiera trunk + FPN neck + prompt-free mask decoder + refinement tail) as one C-ABI call.
// Dataflow of /root/reference/src/sam2_infer.py:220-275 (SAM2ImageWrapper.forward) and of the sam2 package modules
// it calls, re-laid-out for B200:
//   * activations are token-major (channels-last) for the whole path, so every Linear / 1x1 conv / k2s2 transposed
//     conv is one tcgen05 GEMM (gemm_tc.cuh) with the surrounding elementwise work in its epilogue;
//   * residual stream, LayerNorm statistics, softmax and the whole token side of the decoder stay fp32; bf16 is
//     used only for MMA operands;
//   * window partition is fused into LayerNorm (window-major bf16 operand), window un-partition + residual into the
//     proj GEMM epilogue; one block-diagonal flash-attention kernel (attn_tc.cu) serves windowed / pooled / global;
//   * input-independent pieces (positional embeddings, dense prompt, layer-0 token self-attention, PE projections,
//     neck∘conv_s0/s1 products) are folded at load time by the host (circuitvision_b200/sam2_infer.py).
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "mlp_fused.cuh"
#include "patch_embed.cuh"
#include "sam2_kernels.cuh"

namespace cvb {
int attn_tc_launch(const __nv_bfloat16* q, long long ldq, int qcols, int qcol0, const __nv_bfloat16* k, long long ldk,
                   int kcols, int kcol0, const __nv_bfloat16* v, long long ldv, int vcols, int vcol0, int Mq, int Mkv,
                   int Wq, int Wkv, int heads, int D, float scale, __nv_bfloat16* out, long long ld_out, int fp16,
                   cudaStream_t st);
int device_sm_count();
}  // namespace cvb

using namespace cvb;
typedef __nv_bfloat16 bf16;

struct DevTensor {
  void* p = nullptr;
  size_t bytes = 0;
  int dtype = 0;  // 0 f32, 1 bf16
};

struct BlockPlan {
  int dim_in, dim_out, heads, ws, pool;
  int H, W;  // input grid
};

struct cv_sam2 {
  cv_sam2_cfg cfg;
  int device = 0;
  std::map<std::string, DevTensor> w;
  std::map<std::string, DevTensor> buf;
  std::vector<BlockPlan> plan;
  int stage_end[4];
  bool finalized = false;
  int f16 = 0;  // 16-bit operand format: 0 bf16, 1 IEEE half
  int hd = 96;  // real head dim (embed_dim / num_heads: 96 tiny/small, 56 base+, 72 large)
  int D = 96;   // head dim of the attention buffers: hd zero-padded to 64 or 96 by the weight folding
  int launches = 0;
  float refc_b = 0.f;
  int count_sat = 0;  // debug: count fp16 conversions that saturated (|x| == 65504) in every 16-bit activation buffer
};

#define TRY(x)               \
  do {                       \
    int _rc = (x);           \
    if (_rc) return _rc;     \
  } while (0)

static int alloc_buf(cv_sam2* h, const char* name, size_t bytes, int dtype) {
  DevTensor t;
  t.bytes = bytes;
  t.dtype = dtype;
  cudaError_t e = cudaMalloc(&t.p, bytes ? bytes : 16);
  if (e != cudaSuccess) return cvb_fail_cuda(e, name);
  // debug: CVB_SAM2_FILL=<byte> fills every work buffer (or only CVB_SAM2_FILL_ONLY=<name>) at allocation, so that a read of a
  // buffer that was never written shows up whatever the allocator handed out (scripts/sam2_uninit_probe.py)
  if (const char* f = getenv("CVB_SAM2_FILL")) {
    const char* only = getenv("CVB_SAM2_FILL_ONLY");
    if (!only || !strcmp(only, name)) cudaMemset(t.p, atoi(f), bytes ? bytes : 16);
    else cudaMemset(t.p, 0, bytes ? bytes : 16);
  }
  h->buf[name] = t;
  return CV_OK;
}
static const float* WF(cv_sam2* h, const std::string& n) {
  auto it = h->w.find(n);
  return it == h->w.end() ? nullptr : (const float*)it->second.p;
}
static const bf16* WB(cv_sam2* h, const std::string& n) {
  auto it = h->w.find(n);
  return it == h->w.end() ? nullptr : (const bf16*)it->second.p;
}
template <typename T>
static T* BUF(cv_sam2* h, const char* n) { return (T*)h->buf[n].p; }

extern "C" int cv_sam2_create(const cv_sam2_cfg* cfg, int device, cv_sam2** out) {
  if (!cfg || !out) return cvb_fail(CV_ERR_INVALID, "cv_sam2_create: null");
  if (cfg->num_heads <= 0 || cfg->embed_dim % cfg->num_heads || cfg->embed_dim % 16)
    return cvb_fail(CV_ERR_INVALID, "cv_sam2_create: embed_dim must be a multiple of 16 and of num_heads");
  const int hd = cfg->embed_dim / cfg->num_heads;
  if (hd > 96 || hd % 8) return cvb_fail(CV_ERR_INVALID, "cv_sam2_create: head dim must be a multiple of 8 and <= 96");
  cv_sam2* h = new cv_sam2();
  h->hd = hd;
  h->D = hd <= 64 ? 64 : 96;
  h->cfg = *cfg;
  h->device = device;
  h->f16 = cfg->operand_fp16 ? 1 : 0;
  int total = 0, H = 256, W = 256;
  for (int s = 0; s < 4; s++) {
    for (int b = 0; b < cfg->stages[s]; b++) {
      bool first = s > 0 && b == 0;
      BlockPlan p;
      p.dim_out = cfg->embed_dim << s;
      p.dim_in = first ? cfg->embed_dim << (s - 1) : p.dim_out;
      p.heads = cfg->num_heads << s;
      p.ws = first ? cfg->window_spec[s - 1] : cfg->window_spec[s];
      for (int g = 0; g < cfg->n_global; g++)
        if (cfg->global_blocks[g] == total) p.ws = 0;
      p.pool = first ? 1 : 0;
      p.H = H; p.W = W;
      if (p.pool) { H /= 2; W /= 2; }
      h->plan.push_back(p);
      total++;
    }
    h->stage_end[s] = total - 1;
  }
  *out = h;
  return CV_OK;
}

extern "C" int cv_sam2_destroy(cv_sam2* h) {
  if (!h) return CV_OK;
  for (auto& kv : h->w) cudaFree(kv.second.p);
  for (auto& kv : h->buf) cudaFree(kv.second.p);
  delete h;
  return CV_OK;
}

// dtype: 0 = float32, 1 = bfloat16 (raw 16-bit).  host_data is copied to the device.
extern "C" int cv_sam2_set_tensor(cv_sam2* h, const char* name, const void* host_data, int dtype, long long numel) {
  if (!h || !name || !host_data || numel <= 0) return cvb_fail(CV_ERR_INVALID, "cv_sam2_set_tensor: bad argument");
  CVB_CHECK(cudaSetDevice(h->device));
  DevTensor t;
  t.dtype = dtype;
  t.bytes = (size_t)numel * (dtype == 1 ? 2 : 4);
  CVB_CHECK(cudaMalloc(&t.p, t.bytes));
  CVB_CHECK(cudaMemcpy(t.p, host_data, t.bytes, cudaMemcpyHostToDevice));
  auto it = h->w.find(name);
  if (it != h->w.end()) cudaFree(it->second.p);
  h->w[name] = t;
  return CV_OK;
}

static size_t window_rows(const BlockPlan& p, int B) {
  if (p.ws == 0) return (size_t)B * p.H * p.W;
  int nwx = (p.W + p.ws - 1) / p.ws, nwy = (p.H + p.ws - 1) / p.ws;
  return (size_t)B * nwx * nwy * p.ws * p.ws;
}

extern "C" int cv_sam2_finalize(cv_sam2* h) {
  if (!h) return cvb_fail(CV_ERR_INVALID, "cv_sam2_finalize: null");
  CVB_CHECK(cudaSetDevice(h->device));
  const int B = h->cfg.max_batch, E = h->cfg.embed_dim;
  if (B <= 0) return cvb_fail(CV_ERR_INVALID, "cv_sam2_finalize: max_batch");
  h->finalized = false;
  for (auto& kv : h->buf) cudaFree(kv.second.p);  // re-finalize after cv_sam2_set_max_batch
  h->buf.clear();
  // required tensors (fail loudly on an incomplete weight set)
  std::vector<std::string> need = {"pe.w", "pe.b", "pos", "pe.w8", "pos8", "neck3.w", "neck3.b", "neck2.w", "neck2.b", "s1.w", "s1.b", "s0.w",
                                   "s0.b", "dense", "tok0", "l0.q1", "l0.t2i.qc", "up1.w", "up1.b", "upln.g", "upln.b",
                                   "up2.w", "up2.b", "fin.q.w", "fin.kv.w", "fin.kpe", "fin.o.w", "fin.n.g"};
  for (size_t i = 0; i < h->plan.size(); i++) {
    std::string p = "b" + std::to_string(i);
    for (const char* s : {".n1.g", ".n1.b", ".qkv.w", ".qkv.b", ".proj.w", ".proj.b", ".n2.g", ".n2.b", ".fc1.w", ".fc1.b",
                          ".fc2.w", ".fc2.b"})
      need.push_back(p + s);
    if (h->plan[i].dim_in != h->plan[i].dim_out) { need.push_back(p + ".sc.w"); need.push_back(p + ".sc.b"); }
  }
  for (const char* pre : {"l0", "l1"}) {
    std::string p(pre);
    for (const char* s : {".t2i.kv.w", ".t2i.kv.b", ".t2i.kpe", ".t2i.o.w", ".t2i.o.b", ".n2.g", ".n2.b", ".mlp1.w", ".mlp1.b",
                          ".mlp2.w", ".mlp2.b", ".n3.g", ".n3.b", ".i2t.q.w", ".i2t.q.b", ".i2t.qpe", ".i2t.k.w", ".i2t.k.b",
                          ".i2t.v.w", ".i2t.v.b", ".i2t.o.w", ".i2t.o.b", ".n4.g", ".n4.b"})
      need.push_back(p + s);
  }
  for (const char* s : {"l1.sa.q.w", "l1.sa.q.b", "l1.sa.k.w", "l1.sa.k.b", "l1.sa.v.w", "l1.sa.v.b", "l1.sa.o.w", "l1.sa.o.b",
                        "l1.n1.g", "l1.n1.b", "l1.t2i.q.w", "l1.t2i.q.b", "fin.q.b", "fin.kv.b", "fin.o.b", "fin.n.b",
                        "iou.0.w", "iou.0.b", "iou.1.w", "iou.1.b", "iou.2.w", "iou.2.b"})
    need.push_back(s);
  for (int k = 0; k < 4; k++)
    for (int j = 0; j < 3; j++) {
      need.push_back("hyp" + std::to_string(k) + "." + std::to_string(j) + ".w");
      need.push_back("hyp" + std::to_string(k) + "." + std::to_string(j) + ".b");
    }
  if (h->cfg.use_refinement) {
    for (int j = 0; j < 4; j++) { need.push_back("ref" + std::to_string(j) + ".w"); need.push_back("ref" + std::to_string(j) + ".b"); }
    need.push_back("refc.w");
    need.push_back("refc.b");
  }
  for (auto& n : need)
    if (!h->w.count(n)) {
      static thread_local char msg[160];
      snprintf(msg, sizeof(msg), "cv_sam2_finalize: weight tensor '%s' was not provided", n.c_str());
      return cvb_fail(CV_ERR_INVALID, msg);
    }
  size_t maxA = (size_t)B * 65536 * 2 * PE_K, maxQKV = 0, maxAO = 0, maxHd = 0, maxS = 0, maxQp = 0;
  for (auto& p : h->plan) {
    size_t M = window_rows(p, B);
    maxA = std::max(maxA, M * p.dim_in);
    maxA = std::max(maxA, (size_t)B * p.H * p.W * p.dim_out);  // LN2 output / casts
    const size_t Cp = (size_t)p.heads * h->D;  // attention width with the padded head dim
    maxQKV = std::max(maxQKV, M * 3 * Cp);
    size_t Mq = p.pool ? M / 4 : M;
    maxAO = std::max(maxAO, Mq * Cp);
    size_t T = (size_t)B * (p.pool ? (p.H / 2) * (p.W / 2) : p.H * p.W);
    maxHd = std::max(maxHd, T * 4 * p.dim_out);
    if (p.pool) { maxS = std::max(maxS, M * p.dim_out); maxQp = std::max(maxQp, Mq * Cp); }
  }
  TRY(alloc_buf(h, "A", maxA * 2, 1));
  TRY(alloc_buf(h, "QKV", maxQKV * 2, 1));
  TRY(alloc_buf(h, "AO", maxAO * 2, 1));
  TRY(alloc_buf(h, "Hd", maxHd * 2, 1));
  TRY(alloc_buf(h, "S", maxS * 4, 0));
  TRY(alloc_buf(h, "Qp", maxQp * 2, 1));
  for (int s = 0; s < 4; s++) {
    char n[8];
    snprintf(n, sizeof(n), "X%d", s);
    TRY(alloc_buf(h, n, (size_t)B * (65536 >> (2 * s)) * (E << s) * 4, 0));
  }
  for (int s = 0; s < 3; s++) {  // 16-bit copies of the stage outputs (operands of the neck's lateral convs)
    char n[8];
    snprintf(n, sizeof(n), "XS%d", s);
    TRY(alloc_buf(h, n, (size_t)B * (65536 >> (2 * s)) * (E << s) * 2, 1));
  }
  TRY(alloc_buf(h, "s0", (size_t)B * 65536 * 32 * 4, 0));
  TRY(alloc_buf(h, "s1", (size_t)B * 16384 * 64 * 4, 0));
  TRY(alloc_buf(h, "L3", (size_t)B * 1024 * 256 * 4, 0));
  TRY(alloc_buf(h, "keys32", (size_t)B * 4096 * 256 * 4, 0));
  TRY(alloc_buf(h, "keys16", (size_t)B * 4096 * 256 * 2, 1));
  TRY(alloc_buf(h, "KV", (size_t)B * 4096 * 256 * 4, 0));
  TRY(alloc_buf(h, "Qi", (size_t)B * 4096 * 128 * 4, 0));
  TRY(alloc_buf(h, "Ai", (size_t)B * 4096 * 128 * 2, 1));
  const size_t TK = (size_t)B * 38;
  for (const char* n : {"q", "qpe", "t256a", "t256b", "t256c", "t256d"}) TRY(alloc_buf(h, n, TK * 256 * 4, 0));
  TRY(alloc_buf(h, "t128a", TK * 128 * 4, 0));
  TRY(alloc_buf(h, "t128b", TK * 128 * 4, 0));
  TRY(alloc_buf(h, "t2048", TK * 2048 * 4, 0));
  TRY(alloc_buf(h, "t2i_scratch", attn_t2i_scratch_floats(B, 38, 8, 16) * 4, 0));
  TRY(alloc_buf(h, "hyper", (size_t)B * 4 * 32 * 4, 0));
  TRY(alloc_buf(h, "h256a", (size_t)B * 4 * 256 * 4, 0));
  TRY(alloc_buf(h, "h256b", (size_t)B * 4 * 256 * 4, 0));
  TRY(alloc_buf(h, "iou4", (size_t)B * 4 * 4, 0));
  TRY(alloc_buf(h, "U1", (size_t)B * 16384 * 64 * 4, 0));
  TRY(alloc_buf(h, "U1n", (size_t)B * 16384 * 64 * 2, 1));
  TRY(alloc_buf(h, "U2", (size_t)B * 65536 * 32 * 4, 0));
  TRY(alloc_buf(h, "masks", (size_t)B * 4 * 65536 * 4, 0));
  TRY(alloc_buf(h, "counts", (size_t)B * 2 * 4, 0));
  TRY(alloc_buf(h, "low", (size_t)B * 65536 * 4, 0));
  TRY(alloc_buf(h, "iou", (size_t)B * 4, 0));
  TRY(alloc_buf(h, "sel", (size_t)B * 4, 0));
  TRY(alloc_buf(h, "high", (size_t)B * 1024 * 1024 * 4, 0));
  TRY(alloc_buf(h, "satcount", 16, 0));
  CVB_CHECK(cudaMemset(h->buf["satcount"].p, 0, 16));
  if (h->cfg.use_refinement) CVB_CHECK(cudaMemcpy(&h->refc_b, h->w["refc.b"].p, 4, cudaMemcpyDeviceToHost));
  h->finalized = true;
  return CV_OK;
}

// Re-sizes the activation workspace (weights stay resident).
extern "C" int cv_sam2_set_max_batch(cv_sam2* h, int max_batch) {
  if (!h || max_batch <= 0) return cvb_fail(CV_ERR_INVALID, "cv_sam2_set_max_batch: bad argument");
  h->cfg.max_batch = max_batch;
  return cv_sam2_finalize(h);
}

// Debug tap (cv_sam2_set_debug): tc::pack16 converts with cvt.rn.satfinite, so an fp32 value beyond the fp16 range
// becomes +-65504 silently.  When enabled, every 16-bit activation buffer is scanned right after the kernel that wrote it
// and the number of +-65504 entries accumulates in the "satcount" buffer (read with cv_sam2_read_buffer).
static int sat(cv_sam2* h, const void* p, long long n, cudaStream_t st) {
  if (!h->count_sat || !h->f16 || n <= 0) return CV_OK;
  return launch_count_sat16((const uint16_t*)p, n, BUF<unsigned int>(h, "satcount"), st);
}

static int gemm(cv_sam2* h, const bf16* A, long long lda, const bf16* W, int M, int N, int K, GemmEpilogue& e, cudaStream_t st) {
  e.fp16 = h->f16;
  int rc = gemm_tc_launch(A, lda, W, K, M, N, K, e, device_sm_count(), st);
  h->launches++;
  return rc;
}

// one Hiera block; X is the fp32 residual stream of the block's stage (updated in place), Xn the next
// stage's stream for Q-pooled blocks.
static int run_block(cv_sam2* h, int i, int B, float* X, float* Xn, cudaStream_t st, bool ln1_done = false, bf16* x_copy16 = nullptr) {
  const BlockPlan& p = h->plan[i];
  const std::string pre = "b" + std::to_string(i);
  const int ws = p.ws, H = p.H, W = p.W, Cin = p.dim_in, C = p.dim_out;
  const long long T = (long long)B * H * W;
  const int nwx = ws ? (W + ws - 1) / ws : 1, nwy = ws ? (H + ws - 1) / ws : 1;
  const long long M = ws ? (long long)B * nwx * nwy * ws * ws : T;
  const int Wkv = ws ? ws * ws : H * W;
  bf16* A = BUF<bf16>(h, "A");
  bf16* QKV = BUF<bf16>(h, "QKV");
  bf16* AO = BUF<bf16>(h, "AO");
  bf16* Hd = BUF<bf16>(h, "Hd");
  // norm1 (+ window partition with zero pad rows)
  if (!ln1_done) {  // (block 0 after the fused patch embedding: "A" already holds norm1(X) in window-major order)
    // (x_copy16: first block of a stage — X is the finished output of the previous stage, which the neck wants as a 16-bit operand)
    TRY(launch_ln_rows(X, T, Cin, WF(h, pre + ".n1.g"), WF(h, pre + ".n1.b"), 1e-6f, B, H, W, ws, h->f16, A, nullptr, st, x_copy16));
    h->launches++;
  }
  GemmEpilogue e;
  e.bias = WF(h, pre + ".qkv.b");
  e.out_bf16 = QKV;
  const int Cp = p.heads * h->D, D = h->D;  // q | k | v each Cp wide, head i at columns [i*D, i*D + hd) (+ zero padding)
  e.ld_bf16 = 3 * Cp;
  // Q-pooled blocks with windows of 4 / 8: both 2 x 2 max-poolings (q and the shortcut) happen in the GEMM epilogues
  const bool fuse_pool = p.pool && (ws == 4 || ws == 8) && H % ws == 0 && W % ws == 0 && Cp % 32 == 0;
  if (fuse_pool) {
    e.map_mode = GEMM_MAP_QPOOL;
    e.ws = ws;
    e.pool_cols = Cp;
    e.pool_out = BUF<bf16>(h, "Qp");
    e.ld_pool = Cp;
  }
  TRY(gemm(h, A, Cin, WB(h, pre + ".qkv.w"), (int)M, 3 * Cp, Cin, e, st));
  TRY(sat(h, A, M * Cin, st));
  TRY(sat(h, QKV, M * 3 * Cp, st));
  float* Xo = X;
  long long To = T;
  int Ho = H, Wo = W, wso = ws;
  const float scale = 1.0f / sqrtf((float)h->hd);
  if (p.pool) {
    // shortcut = maxpool2x2(proj(norm1(x))), computed window-major then gathered into the new grid
    float* S = BUF<float>(h, "S");
    GemmEpilogue es;
    es.bias = WF(h, pre + ".sc.b");
    es.ld_f32 = C;
    if (fuse_pool) {
      // windows of 4 / 8 tokens per side: the 2 x 2 pooling groups sit inside one warp of the epilogue, the GEMM
      // writes the pooled rows directly (no [M, C] fp32 round trip through HBM)
      es.out_f32 = Xn;
      es.map_mode = GEMM_MAP_POOL2;
      es.ws = ws; es.nwx = nwx; es.nwy = nwy; es.H = H; es.W = W;
      TRY(gemm(h, A, Cin, WB(h, pre + ".sc.w"), (int)M, C, Cin, es, st));
      h->launches--;  // one launch instead of two (the += 2 below counts pool_shortcut + pool_q)
    } else {
      es.out_f32 = S;
      TRY(gemm(h, A, Cin, WB(h, pre + ".sc.w"), (int)M, C, Cin, es, st));
      TRY(launch_pool_shortcut(S, B, H, W, ws, C, Xn, st));
    }
    bf16* Qp = BUF<bf16>(h, "Qp");
    if (!fuse_pool) TRY(launch_pool_q(QKV, 3 * Cp, (int)(M / (ws * ws)), ws, Cp, h->f16, Qp, st));
    h->launches += fuse_pool ? 1 : 2;
    TRY(attn_tc_launch(Qp, Cp, Cp, 0, QKV, 3 * Cp, 3 * Cp, Cp, QKV, 3 * Cp, 3 * Cp, 2 * Cp, (int)(M / 4), (int)M, Wkv / 4, Wkv,
                       p.heads, D, scale, AO, Cp, h->f16, st));
    Xo = Xn; Ho = H / 2; Wo = W / 2; wso = ws / 2; To = T / 4;
  } else {
    TRY(attn_tc_launch(QKV, 3 * Cp, 3 * Cp, 0, QKV, 3 * Cp, 3 * Cp, Cp, QKV, 3 * Cp, 3 * Cp, 2 * Cp, (int)M, (int)M, Wkv, Wkv,
                       p.heads, D, scale, AO, Cp, h->f16, st));
  }
  h->launches++;
  TRY(sat(h, AO, (p.pool ? M / 4 : M) * Cp, st));
  if (p.pool) TRY(sat(h, BUF<bf16>(h, "Qp"), (M / 4) * Cp, st));
  // proj + window un-partition + residual (in place on the residual stream)
  GemmEpilogue ep;
  ep.bias = WF(h, pre + ".proj.b");
  ep.res = Xo; ep.ld_res = C;
  ep.out_f32 = Xo; ep.ld_f32 = C;
  if (wso > 0) {
    ep.map_mode = GEMM_MAP_UNWINDOW;
    ep.ws = wso; ep.nwx = nwx; ep.nwy = nwy; ep.H = Ho; ep.W = Wo;
  }
  TRY(gemm(h, AO, Cp, WB(h, pre + ".proj.w"), (int)(p.pool ? M / 4 : M), C, Cp, ep, st));
  // norm2 -> MLP (GELU) -> residual: one fused kernel for the stage-1/2 widths (normalised operand, hidden activation and
  // fc2 accumulator stay on the SM), LayerNorm + two GEMMs otherwise
  if (mlp_fused_supported(C)) {
    MlpFusedArgs m;
    m.X = Xo; m.M = (int)To; m.C = C;
    m.gamma = WF(h, pre + ".n2.g"); m.beta = WF(h, pre + ".n2.b"); m.eps = 1e-6f;
    m.W1 = WB(h, pre + ".fc1.w"); m.b1 = WF(h, pre + ".fc1.b");
    m.W2 = WB(h, pre + ".fc2.w"); m.b2 = WF(h, pre + ".fc2.b");
    m.fp16 = h->f16;
    m.sat_counter = h->count_sat ? BUF<unsigned int>(h, "satcount") : nullptr;
    TRY(mlp_fused_launch(m, device_sm_count(), st));
    h->launches++;
    return CV_OK;
  }
  TRY(launch_ln_rows(Xo, To, C, WF(h, pre + ".n2.g"), WF(h, pre + ".n2.b"), 1e-6f, B, Ho, Wo, 0, h->f16, A, nullptr, st));
  h->launches++;
  GemmEpilogue e1;
  e1.bias = WF(h, pre + ".fc1.b");
  e1.act = GEMM_ACT_GELU;
  e1.out_bf16 = Hd; e1.ld_bf16 = 4 * C;
  TRY(gemm(h, A, C, WB(h, pre + ".fc1.w"), (int)To, 4 * C, C, e1, st));
  TRY(sat(h, A, To * C, st));
  TRY(sat(h, Hd, To * 4 * C, st));
  GemmEpilogue e2;
  e2.bias = WF(h, pre + ".fc2.b");
  e2.res = Xo; e2.ld_res = C;
  e2.out_f32 = Xo; e2.ld_f32 = C;
  TRY(gemm(h, Hd, 4 * C, WB(h, pre + ".fc2.w"), (int)To, C, 4 * C, e2, st));
  return CV_OK;
}

static int tok_lin(cv_sam2* h, const float* A, long long lda, const std::string& name, int R, int N, int K, int act,
                   const float* res, float* C, long long ldc, cudaStream_t st) {
  const float* W = WF(h, name + ".w");
  if (!W) {
    static thread_local char msg[160];
    snprintf(msg, sizeof(msg), "cv_sam2_forward: weight tensor '%s.w' missing", name.c_str());
    return cvb_fail(CV_ERR_INVALID, msg);
  }
  h->launches++;
  return launch_tok_linear(A, lda, W, WF(h, name + ".b"), R, N, K, act, res, N, C, ldc, st);
}

// tokens -> image cross attention + residual + LayerNorm (layers 0/1 and the final attention)
static int t2i_block(cv_sam2* h, const std::string& pre, const std::string& norm, int B, bool const_q, cudaStream_t st) {
  const int T = 38, R = B * T;
  float* q = BUF<float>(h, "q");
  float* KV = BUF<float>(h, "KV");
  GemmEpilogue e;
  e.bias = WF(h, pre + ".kv.b");
  e.res = WF(h, pre + ".kpe"); e.ld_res = 256; e.res_row_mod = 4096;
  e.out_f32 = KV; e.ld_f32 = 256;
  TRY(gemm(h, BUF<bf16>(h, "keys16"), 256, WB(h, pre + ".kv.w"), B * 4096, 256, 256, e, st));
  float* att = BUF<float>(h, "t128a");
  const float* qproj;
  long long qstride;
  if (const_q) {
    qproj = WF(h, pre + ".qc");
    qstride = 0;
  } else {
    TRY(launch_tok_add_bcast(q, WF(h, "tok0"), T, R, 256, BUF<float>(h, "qpe"), st));
    TRY(tok_lin(h, BUF<float>(h, "qpe"), 256, pre + ".q", R, 128, 256, 0, nullptr, BUF<float>(h, "t128b"), 128, st));
    qproj = BUF<float>(h, "t128b");
    qstride = (long long)T * 128;
    h->launches++;
  }
  TRY(launch_attn_t2i(qproj, qstride, KV, KV + 128, 256, B, T, 4096, 8, 16, att, BUF<float>(h, "t2i_scratch"), st));
  h->launches += 2;
  float* o = BUF<float>(h, "t256a");
  TRY(tok_lin(h, att, 128, pre + ".o", R, 256, 128, 0, nullptr, o, 256, st));
  // queries = LN(queries + attn_out); in layer 0 the incoming queries are the load-time constant l0.q1
  const float* add = const_q ? WF(h, "l0.q1") : q;
  TRY(launch_tok_add_ln(o, add, const_q ? T : 0, WF(h, norm + ".g"), WF(h, norm + ".b"), 1e-5f, R, 256, q, st));
  h->launches++;
  return CV_OK;
}

static int decoder_layer(cv_sam2* h, int l, int B, cudaStream_t st) {
  const int T = 38, R = B * T;
  const std::string pre = "l" + std::to_string(l);
  float* q = BUF<float>(h, "q");
  float* qpe = BUF<float>(h, "qpe");
  const float* tok0 = WF(h, "tok0");
  if (l > 0) {
    // self attention: q = k = queries + pe, v = queries; queries = LN1(queries + out)
    TRY(launch_tok_add_bcast(q, tok0, T, R, 256, qpe, st));
    TRY(tok_lin(h, qpe, 256, pre + ".sa.q", R, 256, 256, 0, nullptr, BUF<float>(h, "t256a"), 256, st));
    TRY(tok_lin(h, qpe, 256, pre + ".sa.k", R, 256, 256, 0, nullptr, BUF<float>(h, "t256b"), 256, st));
    TRY(tok_lin(h, q, 256, pre + ".sa.v", R, 256, 256, 0, nullptr, BUF<float>(h, "t256c"), 256, st));
    TRY(launch_tok_self_attn(BUF<float>(h, "t256a"), BUF<float>(h, "t256b"), BUF<float>(h, "t256c"), B, T, 8, 32,
                             BUF<float>(h, "t256d"), st));
    TRY(tok_lin(h, BUF<float>(h, "t256d"), 256, pre + ".sa.o", R, 256, 256, 0, nullptr, BUF<float>(h, "t256a"), 256, st));
    TRY(launch_tok_add_ln(BUF<float>(h, "t256a"), q, 0, WF(h, pre + ".n1.g"), WF(h, pre + ".n1.b"), 1e-5f, R, 256, q, st));
    h->launches += 3;
  }
  TRY(t2i_block(h, pre + ".t2i", pre + ".n2", B, l == 0, st));
  // MLP (ReLU) + LN3
  TRY(tok_lin(h, q, 256, pre + ".mlp1", R, 2048, 256, 2, nullptr, BUF<float>(h, "t2048"), 2048, st));
  TRY(tok_lin(h, BUF<float>(h, "t2048"), 2048, pre + ".mlp2", R, 256, 2048, 0, nullptr, BUF<float>(h, "t256a"), 256, st));
  TRY(launch_tok_add_ln(BUF<float>(h, "t256a"), q, 0, WF(h, pre + ".n3.g"), WF(h, pre + ".n3.b"), 1e-5f, R, 256, q, st));
  h->launches++;
  // image -> tokens: q = keys + key_pe, k = queries + pe, v = queries; keys = LN4(keys + out)
  float* Qi = BUF<float>(h, "Qi");
  GemmEpilogue e;
  e.bias = WF(h, pre + ".i2t.q.b");
  e.res = WF(h, pre + ".i2t.qpe"); e.ld_res = 128; e.res_row_mod = 4096;
  e.out_f32 = Qi; e.ld_f32 = 128;
  TRY(gemm(h, BUF<bf16>(h, "keys16"), 256, WB(h, pre + ".i2t.q.w"), B * 4096, 128, 256, e, st));
  TRY(launch_tok_add_bcast(q, tok0, T, R, 256, qpe, st));
  TRY(tok_lin(h, qpe, 256, pre + ".i2t.k", R, 128, 256, 0, nullptr, BUF<float>(h, "t128a"), 128, st));
  TRY(tok_lin(h, q, 256, pre + ".i2t.v", R, 128, 256, 0, nullptr, BUF<float>(h, "t128b"), 128, st));
  TRY(launch_attn_i2t(Qi, 128, BUF<float>(h, "t128a"), BUF<float>(h, "t128b"), B, 4096, T, 8, 16, h->f16, BUF<bf16>(h, "Ai"), st));
  TRY(sat(h, BUF<bf16>(h, "Ai"), (long long)B * 4096 * 128, st));
  h->launches += 2;
  float* keys32 = BUF<float>(h, "keys32");
  GemmEpilogue eo;
  eo.bias = WF(h, pre + ".i2t.o.b");
  eo.res = keys32; eo.ld_res = 256;
  eo.out_f32 = keys32; eo.ld_f32 = 256;
  TRY(gemm(h, BUF<bf16>(h, "Ai"), 128, WB(h, pre + ".i2t.o.w"), B * 4096, 256, 128, eo, st));
  TRY(launch_ln_rows(keys32, (long long)B * 4096, 256, WF(h, pre + ".n4.g"), WF(h, pre + ".n4.b"), 1e-5f, B, 64, 64, 0,
                     h->f16, BUF<bf16>(h, "keys16"), keys32, st));
  h->launches++;
  TRY(sat(h, BUF<bf16>(h, "keys16"), (long long)B * 4096 * 256, st));
  return CV_OK;
}

// images: input_kind 0 = uint8 HWC [B,1024,1024,3] (ToTensor + Normalize fused into the patch gather; swap_rb applies
// the BGR<->RGB swap of circuit_analyzer.py:343), 1 = float32 CHW [B,3,1024,1024] already normalised.
extern "C" int cv_sam2_forward(cv_sam2* h, const void* images, int input_kind, int swap_rb, int B, float* low_res,
                               float* iou, float* high_res, uint8_t* mask_u8, int out_h, int out_w, float* out_logits,
                               int* extents, void* stream) {
  cvb_reset_launches();
  if (!h || !h->finalized) return cvb_fail(CV_ERR_INVALID, "cv_sam2_forward: engine not finalized");
  if (!images || B <= 0 || B > h->cfg.max_batch) return cvb_fail(CV_ERR_INVALID, "cv_sam2_forward: bad batch");
  if ((mask_u8 || out_logits) && (out_h <= 0 || out_w <= 0)) return cvb_fail(CV_ERR_INVALID, "cv_sam2_forward: output size");
  cudaStream_t st = (cudaStream_t)stream;
  h->launches = 0;
  if (h->count_sat) CVB_CHECK(cudaMemsetAsync(BUF<unsigned int>(h, "satcount"), 0, 16, st));
  const int E = h->cfg.embed_dim;
  static const float mean[3] = {0.485f, 0.456f, 0.406f};
  static const float istd[3] = {1.0f / 0.229f, 1.0f / 0.224f, 1.0f / 0.225f};
  bf16* A = BUF<bf16>(h, "A");
  float* X[4] = {BUF<float>(h, "X0"), BUF<float>(h, "X1"), BUF<float>(h, "X2"), BUF<float>(h, "X3")};
  // ---- patch embed (two-term bf16 split of the pixels) + positional embedding
  (void)mean; (void)istd;
  bool ln1_fused = false;
  if (input_kind == 0) {
    // raw pixels (exact in bf16); 1/255, mean/std, the conv bias and the positional embedding live in pe.w8 / pos8
    if (patch_embed_supported(E, 1024)) {
      // pixels -> conv -> + pos -> X0 and, when block 0 is an 8 x 8 windowed block of the same width, its norm1 output as well
      const BlockPlan& p0 = h->plan[0];
      static const int pe_ln_on = getenv("CVB_PATCH_LN") ? atoi(getenv("CVB_PATCH_LN")) : 1;  // A/B switch
      ln1_fused = pe_ln_on && !p0.pool && p0.ws == 8 && p0.dim_in == E && p0.H == 256 && p0.W == 256;
      PatchEmbedArgs pa;
      pa.img = (const uint8_t*)images; pa.B = B; pa.S = 1024; pa.swap_rb = swap_rb; pa.fp16 = h->f16;
      pa.W = WB(h, "pe.w8"); pa.E = E; pa.pos = WF(h, "pos8"); pa.X0 = X[0];
      pa.gamma = ln1_fused ? WF(h, "b0.n1.g") : nullptr;
      pa.beta = ln1_fused ? WF(h, "b0.n1.b") : nullptr;
      pa.eps = 1e-6f;
      pa.A16 = ln1_fused ? A : nullptr;
      TRY(patch_embed_launch(pa, device_sm_count(), st));
    } else {
      TRY(launch_im2col_u8raw((const uint8_t*)images, B, 1024, swap_rb, h->f16, A, st));
      GemmEpilogue e;
      e.res = WF(h, "pos8"); e.ld_res = E; e.res_row_mod = 65536;
      e.out_f32 = X[0]; e.ld_f32 = E;
      TRY(gemm(h, A, PE_K8, WB(h, "pe.w8"), B * 65536, E, PE_K8, e, st));
    }
  } else {
    TRY(launch_im2col_f32((const float*)images, B, 1024, h->f16, A, st));
    GemmEpilogue e;
    e.bias = WF(h, "pe.b");
    e.res = WF(h, "pos"); e.ld_res = E; e.res_row_mod = 65536;
    e.out_f32 = X[0]; e.ld_f32 = E;
    TRY(gemm(h, A, 2 * PE_K, WB(h, "pe.w"), B * 65536, E, 2 * PE_K, e, st));
  }
  h->launches++;
  if (getenv("CVB_SAM2_STOP_AFTER_PE")) return CV_OK;  // debug (scripts/sam2_determinism_probe.py): X0 / A as the patch embedding left them
  // ---- trunk
  int stage = 0;
  bool have_copy[3] = {false, false, false};
  for (size_t i = 0; i < h->plan.size(); i++) {
    if (h->plan[i].pool) stage++;
    bf16* copy16 = nullptr;
    static const int copy_on = getenv("CVB_LN_COPY") ? atoi(getenv("CVB_LN_COPY")) : 1;  // A/B switch
    if (copy_on && h->plan[i].pool && stage >= 1 && stage <= 3) {
      copy16 = BUF<bf16>(h, stage == 1 ? "XS0" : stage == 2 ? "XS1" : "XS2");
      have_copy[stage - 1] = true;
    }
    TRY(run_block(h, (int)i, B, h->plan[i].pool ? X[stage - 1] : X[stage], X[stage], st, i == 0 && ln1_fused, copy16));
  }
  // ---- neck (level 3 lateral, level 2 lateral + top-down + dense prompt) and the folded conv_s0 / conv_s1
  float* keys32 = BUF<float>(h, "keys32");
  {
    TRY(launch_ln_rows(X[3], (long long)B * 1024, 8 * E, nullptr, nullptr, 0.f, B, 32, 32, 0, h->f16, A, nullptr, st));
    GemmEpilogue e;
    e.bias = WF(h, "neck3.b");
    e.out_f32 = BUF<float>(h, "L3"); e.ld_f32 = 256;
    TRY(gemm(h, A, 8 * E, WB(h, "neck3.w"), B * 1024, 256, 8 * E, e, st));
    const bf16* A2 = have_copy[2] ? BUF<bf16>(h, "XS2") : A;
    if (!have_copy[2]) TRY(launch_ln_rows(X[2], (long long)B * 4096, 4 * E, nullptr, nullptr, 0.f, B, 64, 64, 0, h->f16, A, nullptr, st));
    GemmEpilogue e2;
    e2.bias = WF(h, "neck2.b");
    e2.res = WF(h, "dense"); e2.ld_res = 256; e2.res_row_mod = 4096;  // src = image_embed + dense prompt
    e2.out_f32 = keys32; e2.ld_f32 = 256;
    TRY(gemm(h, A2, 4 * E, WB(h, "neck2.w"), B * 4096, 256, 4 * E, e2, st));
    TRY(launch_add_nearest2(keys32, BUF<float>(h, "L3"), B, 64, 64, 256, st));
    TRY(launch_ln_rows(keys32, (long long)B * 4096, 256, nullptr, nullptr, 0.f, B, 64, 64, 0, h->f16, BUF<bf16>(h, "keys16"), nullptr, st));
    TRY(sat(h, BUF<bf16>(h, "keys16"), (long long)B * 4096 * 256, st));
    const bf16* A1 = have_copy[1] ? BUF<bf16>(h, "XS1") : A;
    if (!have_copy[1]) TRY(launch_ln_rows(X[1], (long long)B * 16384, 2 * E, nullptr, nullptr, 0.f, B, 128, 128, 0, h->f16, A, nullptr, st));
    GemmEpilogue e3;
    e3.bias = WF(h, "s1.b");
    e3.out_f32 = BUF<float>(h, "s1"); e3.ld_f32 = 64;
    TRY(gemm(h, A1, 2 * E, WB(h, "s1.w"), B * 16384, 64, 2 * E, e3, st));
    const bf16* A0 = have_copy[0] ? BUF<bf16>(h, "XS0") : A;
    if (!have_copy[0]) TRY(launch_ln_rows(X[0], (long long)B * 65536, E, nullptr, nullptr, 0.f, B, 256, 256, 0, h->f16, A, nullptr, st));
    GemmEpilogue e4;
    e4.bias = WF(h, "s0.b");
    e4.out_f32 = BUF<float>(h, "s0"); e4.ld_f32 = 32;
    TRY(gemm(h, A0, E, WB(h, "s0.w"), B * 65536, 32, E, e4, st));
    h->launches += 6 - (int)have_copy[0] - (int)have_copy[1] - (int)have_copy[2];
  }
  // ---- mask decoder
  TRY(decoder_layer(h, 0, B, st));
  TRY(decoder_layer(h, 1, B, st));
  TRY(t2i_block(h, "fin", "fin.n", B, false, st));
  float* q = BUF<float>(h, "q");
  {
    // IoU head on token 1, hyper-network MLPs on tokens 2..5 (rows of q with pitch 38*256)
    const long long ldq = 38 * 256;
    float* a = BUF<float>(h, "h256a");
    float* b2 = BUF<float>(h, "h256b");
    TRY(tok_lin(h, q + 256, ldq, "iou.0", B, 256, 256, 2, nullptr, a, 256, st));
    TRY(tok_lin(h, a, 256, "iou.1", B, 256, 256, 2, nullptr, b2, 256, st));
    TRY(tok_lin(h, b2, 256, "iou.2", B, 4, 256, 3, nullptr, BUF<float>(h, "iou4"), 4, st));
    for (int k = 0; k < 4; k++) {
      std::string n = "hyp" + std::to_string(k);
      TRY(tok_lin(h, q + (2 + k) * 256, ldq, n + ".0", B, 256, 256, 2, nullptr, a, 256, st));
      TRY(tok_lin(h, a, 256, n + ".1", B, 256, 256, 2, nullptr, b2, 256, st));
      TRY(tok_lin(h, b2, 256, n + ".2", B, 32, 256, 0, nullptr, BUF<float>(h, "hyper") + k * 32, 128, st));
    }
  }
  {
    // upscaling: ConvT(256->64,k2s2) + s1 -> LayerNorm2d -> GELU -> ConvT(64->32,k2s2) + s0 -> GELU
    GemmEpilogue e;
    e.bias = WF(h, "up1.b");
    e.res = BUF<float>(h, "s1"); e.ld_res = 64;
    e.out_f32 = BUF<float>(h, "U1"); e.ld_f32 = 64;
    e.map_mode = GEMM_MAP_SHUFFLE2; e.H = 64; e.W = 64; e.cout = 64;
    TRY(gemm(h, BUF<bf16>(h, "keys16"), 256, WB(h, "up1.w"), B * 4096, 256, 256, e, st));
    TRY(launch_ln2d_gelu(BUF<float>(h, "U1"), (long long)B * 16384, 64, WF(h, "upln.g"), WF(h, "upln.b"), 1e-6f, h->f16,
                         BUF<bf16>(h, "U1n"), st));
    TRY(sat(h, BUF<bf16>(h, "U1n"), (long long)B * 16384 * 64, st));
    GemmEpilogue e2;
    e2.bias = WF(h, "up2.b");
    e2.res = BUF<float>(h, "s0"); e2.ld_res = 32;
    e2.res_before_act = 1; e2.act = GEMM_ACT_GELU;
    e2.out_f32 = BUF<float>(h, "U2"); e2.ld_f32 = 32;
    e2.map_mode = GEMM_MAP_SHUFFLE2; e2.H = 128; e2.W = 128; e2.cout = 32;
    TRY(gemm(h, BUF<bf16>(h, "U1n"), 64, WB(h, "up2.w"), B * 16384, 128, 64, e2, st));
    h->launches++;
  }
  float* low = BUF<float>(h, "low");
  TRY(launch_mask_product(BUF<float>(h, "U2"), BUF<float>(h, "hyper"), B, 65536, 0.05f, BUF<float>(h, "masks"),
                          BUF<unsigned int>(h, "counts"), st));
  TRY(launch_select_mask(BUF<float>(h, "masks"), BUF<float>(h, "iou4"), BUF<unsigned int>(h, "counts"), B, 65536, 0.98f, low,
                         BUF<float>(h, "iou"), BUF<int>(h, "sel"), st));
  h->launches += 2;
  if (low_res) CVB_CHECK(cudaMemcpyAsync(low_res, low, (size_t)B * 65536 * 4, cudaMemcpyDeviceToDevice, st));
  if (iou) CVB_CHECK(cudaMemcpyAsync(iou, BUF<float>(h, "iou"), (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
  // ---- tail: x4 bilinear + refinement (+ threshold / resize to the caller's size)
  if (high_res || mask_u8 || out_logits) {
    RefineWeights rw;
    rw.use_refine = h->cfg.use_refinement;
    if (rw.use_refine) {
      for (int j = 0; j < 4; j++) {
        rw.w[j] = WF(h, "ref" + std::to_string(j) + ".w");
        rw.b[j] = WF(h, "ref" + std::to_string(j) + ".b");
        if (!rw.w[j] || !rw.b[j]) return cvb_fail(CV_ERR_INVALID, "cv_sam2_forward: refinement weights missing");
      }
      rw.cw = WF(h, "refc.w");
      rw.cb = h->refc_b;
      rw.comp = h->w.count("ref.comp") ? WF(h, "ref.comp") : nullptr;
    } else {
      for (int j = 0; j < 4; j++) { rw.w[j] = nullptr; rw.b[j] = nullptr; }
      rw.cw = nullptr;
      rw.cb = 0.f;
      rw.comp = nullptr;
    }
    const bool native = (out_h == 1024 && out_w == 1024);
    const bool need_resize = (mask_u8 || out_logits) && !native;
    float* hi = high_res ? high_res : ((need_resize || (out_logits && native)) ? BUF<float>(h, "high") : nullptr);
    if (out_logits && native && !high_res) hi = out_logits;
    TRY(launch_tail(low, 0, B, rw, hi, native ? mask_u8 : nullptr, native ? extents : nullptr, st));
    h->launches += 1 + (extents && native ? 1 : 0) + (rw.use_refine && rw.comp ? 1 : 0);  // k_tail (+ k_init_extents) (+ k_tail_phase)
    if (out_logits && native && high_res)
      CVB_CHECK(cudaMemcpyAsync(out_logits, high_res, (size_t)B * 1024 * 1024 * 4, cudaMemcpyDeviceToDevice, st));
    if (need_resize) {
      TRY(launch_resize_threshold(hi, B, 1024, out_h, out_w, out_logits, mask_u8, extents, st));
      h->launches += 1 + (extents ? 1 : 0);
    }
  }
  return CV_OK;
}

extern "C" int cv_sam2_last_launches(cv_sam2* h) { return h ? h->launches : 0; }

extern "C" int cv_sam2_set_debug(cv_sam2* h, int count_fp16_saturation) {
  if (!h) return cvb_fail(CV_ERR_INVALID, "cv_sam2_set_debug: null");
  h->count_sat = count_fp16_saturation ? 1 : 0;
  return CV_OK;
}

// Copies an internal activation buffer (debug / parity taps) into caller memory on the device.
extern "C" int cv_sam2_read_buffer(cv_sam2* h, const char* name, void* dst_device, long long bytes, void* stream) {
  if (!h || !name || !dst_device) return cvb_fail(CV_ERR_INVALID, "cv_sam2_read_buffer: null");
  auto it = h->buf.find(name);
  if (it == h->buf.end()) return cvb_fail(CV_ERR_INVALID, "cv_sam2_read_buffer: unknown buffer");
  if ((size_t)bytes > it->second.bytes) return cvb_fail(CV_ERR_INVALID, "cv_sam2_read_buffer: larger than the buffer");
  CVB_CHECK(cudaMemcpyAsync(dst_device, it->second.p, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return CV_OK;
}

// SAM2Transforms.__call__ (sam2_infer.py:49-51) for one uint8 HWC image of any size: float CHW [3,1024,1024] normalised.
// tmp: device scratch of H*1024*3 floats.
extern "C" int cv_sam2_preprocess(const uint8_t* img_hwc, int H, int W, int swap_rb, float* tmp, float* out_chw, void* stream) {
  cvb_reset_launches();
  if (!img_hwc || !tmp || !out_chw || H <= 0 || W <= 0) return cvb_fail(CV_ERR_INVALID, "cv_sam2_preprocess: bad argument");
  static const float mean[3] = {0.485f, 0.456f, 0.406f};
  static const float istd[3] = {1.0f / 0.229f, 1.0f / 0.224f, 1.0f / 0.225f};
  return launch_preprocess_aa(img_hwc, H, W, 1024, mean, istd, swap_rb, tmp, out_chw, (cudaStream_t)stream);
}

extern "C" int cv_sam2_preprocess_pages(const uint8_t* pages, const void* geom, int B, int max_crop_h, int swap_rb, float* tmp,
                                        float* out_chw, void* stream) {
  cvb_reset_launches();
  if (!pages || !geom || !tmp || !out_chw || B <= 0 || max_crop_h <= 0)
    return cvb_fail(CV_ERR_INVALID, "cv_sam2_preprocess_pages: bad argument");
  static const float mean[3] = {0.485f, 0.456f, 0.406f};
  static const float istd[3] = {1.0f / 0.229f, 1.0f / 0.224f, 1.0f / 0.225f};
  return launch_preprocess_pages(pages, geom, B, max_crop_h, 1024, mean, istd, swap_rb, tmp, out_chw, (cudaStream_t)stream);
}

extern "C" int cv_sam2_resize_logits(const float* logits, int B, int S, int H, int W, float* out_logits, uint8_t* mask_u8,
                                     int* extents, void* stream) {
  cvb_reset_launches();
  if (!logits || B <= 0 || S <= 0 || H <= 0 || W <= 0 || (!out_logits && !mask_u8))
    return cvb_fail(CV_ERR_INVALID, "cv_sam2_resize_logits: bad argument");
  return launch_resize_threshold(logits, B, S, H, W, out_logits, mask_u8, extents, (cudaStream_t)stream);
}

extern "C" int cv_sam2_refine(const float* x, int B, const float* const* w, const float* const* b, const float* cw, float cb,
                              float* out, void* stream) {
  cvb_reset_launches();
  if (!x || !w || !b || !cw || !out || B <= 0) return cvb_fail(CV_ERR_INVALID, "cv_sam2_refine: bad argument");
  RefineWeights rw;
  for (int j = 0; j < 4; j++) { rw.w[j] = w[j]; rw.b[j] = b[j]; }
  rw.cw = cw;
  rw.cb = cb;
  rw.use_refine = 1;
  rw.comp = nullptr;
  return launch_tail(x, 1, B, rw, out, nullptr, nullptr, (cudaStream_t)stream);
}
