// C ABI of the LEAF attack engine (include/leaf_b200.h). Host-side orchestration only: every computation is
// a CUDA kernel launched from here (k1_tokenize.cuh, gemm_sm100.cuh, tower_kernels.cuh). No CPU compute path.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/leaf_b200.h"
#include "gemm2_sm100.cuh"
#include "gemm_sm100.cuh"
#include "k1_tables_host.h"
#include "k1_tokenize.cuh"
#include "tower_kernels.cuh"

using namespace leaf;

static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CK(expr)                                                                                       \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess) return fail(LEAF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                       __FILE__, __LINE__);                                            \
  } while (0)

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct LayerW {
  __nv_bfloat16 *qkv_w = nullptr, *out_w = nullptr, *fc1_w = nullptr, *fc2_w = nullptr;   // engine-owned bf16 copies
  float* qkv_b_own = nullptr;                                                              // HF layout only
  const float* qkv_b = nullptr;
};

struct leaf_engine {
  leaf_cfg_t cfg;
  int device = 0;
  int sm_count = 148;
  encode_tiled_fn encode_tiled = nullptr;
  // K1 tables
  bool bpe_loaded = false;
  std::vector<void*> table_allocs;
  K1Tables tables{};
  // weights
  bool bound = false;
  leaf_weight_ptrs_t wp{};
  std::vector<leaf_layer_ptrs_t> layer_ptrs;
  std::vector<LayerW> lw;
  __nv_bfloat16* proj_w = nullptr;    // [E, W] bf16
  // workspace
  int max_seqs = 0;
  long rows_cap = 0;
  float* x = nullptr;                 // [rows_cap, W] fp32 residual stream
  __nv_bfloat16* h = nullptr;         // [rows_cap, W]  LN output / attention output
  __nv_bfloat16* big = nullptr;       // [rows_cap, 4W] qkv (3W) or MLP hidden (4W)
  __nv_bfloat16* pooled = nullptr;    // [max_seqs, W]
  int *cu = nullptr, *eos_row = nullptr, *total_rows = nullptr, *pfx = nullptr, *own_len = nullptr;
  int4* meta = nullptr;
  // bookkeeping
  int64_t launches = 0;
  bool timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> gemm_events;
  std::vector<cudaEvent_t> event_pool;
  std::map<std::tuple<const void*, long, long, int>, CUtensorMap> tmaps;
};

extern "C" const char* leaf_last_error(void) { return g_err.c_str(); }
extern "C" const char* leaf_version(void) { return "leaf_b200 0.1 (sm_100a)"; }

static int make_tmap(leaf_engine* e, const void* ptr, long rows, long cols, int box_rows, CUtensorMap* out) {
  auto key = std::make_tuple(ptr, rows, cols, box_rows);
  auto it = e->tmaps.find(key);
  if (it != e->tmaps.end()) { *out = it->second; return LEAF_OK; }
  if (cols % 8 != 0) return fail(LEAF_ERR_INVALID, "GEMM K (%ld) must be a multiple of 8", cols);
  if (reinterpret_cast<uintptr_t>(ptr) & 15) return fail(LEAF_ERR_INVALID, "GEMM operand must be 16-byte aligned");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(GEMM_BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = e->encode_tiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(LEAF_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%ld cols=%ld", (int)r, rows, cols);
  if (e->tmaps.size() > 4096) e->tmaps.clear();
  e->tmaps[key] = m;
  *out = m;
  return LEAF_OK;
}

static cudaEvent_t get_event(leaf_engine* e) {
  if (!e->event_pool.empty()) { cudaEvent_t ev = e->event_pool.back(); e->event_pool.pop_back(); return ev; }
  cudaEvent_t ev;
  cudaEventCreate(&ev);
  return ev;
}

// C[M,N] = A[M,K] . Bt[N,K]^T with the fused epilogue `epi`
static int launch_gemm(leaf_engine* e, const __nv_bfloat16* A, long a_rows, const __nv_bfloat16* Bt, const float* bias,
                       void* C, int ldc, int M, int N, int K, int epi, int act, const int* m_dev, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return fail(LEAF_ERR_INVALID, "GEMM shape %dx%dx%d", M, N, K);
  if (N % 8 != 0) return fail(LEAF_ERR_INVALID, "GEMM N (%d) must be a multiple of 8", N);
  CUtensorMap ta, tb;
  const int a_box = a_rows < 128 ? static_cast<int>(a_rows) : 128;
  const int b_box = N < 128 ? N : 128;
  int rc = make_tmap(e, A, a_rows, K, a_box, &ta);
  if (rc) return rc;
  rc = make_tmap(e, Bt, N, K, b_box, &tb);
  if (rc) return rc;
  GemmParams p;
  p.tx_bytes = static_cast<uint32_t>(a_box + b_box) * GEMM_BK * 2 * 2;       // both CTAs of the pair
  p.M = M; p.m_dev = m_dev; p.N = N; p.K = K; p.bias = bias; p.C = C; p.ldc = ldc; p.act = act;
  const int m_tiles = (M + GEMM2_BM - 1) / GEMM2_BM, n_tiles = (N + GEMM_BN - 1) / GEMM_BN;
  const long tiles = static_cast<long>(m_tiles) * n_tiles;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (e->timing) { e0 = get_event(e); e1 = get_event(e); cudaEventRecord(e0, st); }
  const int pairs_max = e->sm_count / 2;
  const int pairs = static_cast<int>(tiles < pairs_max ? tiles : pairs_max);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = GEMM2_SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  switch (epi) {
    case EPI_BF16: CK(cudaLaunchKernelEx(&cfg, gemm2_bf16_tn_kernel<EPI_BF16>, ta, tb, p)); break;
    case EPI_BF16_ACT: CK(cudaLaunchKernelEx(&cfg, gemm2_bf16_tn_kernel<EPI_BF16_ACT>, ta, tb, p)); break;
    case EPI_F32_RESIDUAL: CK(cudaLaunchKernelEx(&cfg, gemm2_bf16_tn_kernel<EPI_F32_RESIDUAL>, ta, tb, p)); break;
    case EPI_F32: CK(cudaLaunchKernelEx(&cfg, gemm2_bf16_tn_kernel<EPI_F32>, ta, tb, p)); break;
    default: return fail(LEAF_ERR_INVALID, "unknown epilogue %d", epi);
  }
  if (e->timing) { cudaEventRecord(e1, st); e->gemm_events.emplace_back(e0, e1); }
  e->launches++;
  CK(cudaGetLastError());
  return LEAF_OK;
}

extern "C" int leaf_create(const leaf_cfg_t* cfg, leaf_handle_t* out) {
  if (!cfg || !out) return fail(LEAF_ERR_INVALID, "null argument");
  if (cfg->width <= 0 || cfg->width % 128 != 0 || cfg->width > 2048)
    return fail(LEAF_ERR_INVALID, "width %d must be a multiple of 128 in (0, 2048]", cfg->width);
  if (cfg->heads <= 0 || cfg->width != cfg->heads * 64) return fail(LEAF_ERR_INVALID, "head_dim must be 64 (W=%d, H=%d)", cfg->width, cfg->heads);
  if (cfg->embed_dim <= 0 || cfg->embed_dim % 8 != 0) return fail(LEAF_ERR_INVALID, "embed_dim %d must be a multiple of 8", cfg->embed_dim);
  if (cfg->layers <= 0) return fail(LEAF_ERR_INVALID, "layers");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(LEAF_ERR_CUDA, "no CUDA device: leaf_b200 has no CPU path");
  int dev = 0;
  CK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail(LEAF_ERR_CUDA, "device %s is sm_%d%d; leaf_b200 is built for sm_100a only", prop.name, prop.major, prop.minor);
  leaf_engine* e = new leaf_engine();
  e->cfg = *cfg;
  e->device = dev;
  e->sm_count = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
    delete e;
    return fail(LEAF_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  }
  e->encode_tiled = reinterpret_cast<encode_tiled_fn>(fn);
  CK(cudaFuncSetAttribute(gemm2_bf16_tn_kernel<EPI_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM2_SMEM_BYTES));
  CK(cudaFuncSetAttribute(gemm2_bf16_tn_kernel<EPI_BF16_ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM2_SMEM_BYTES));
  CK(cudaFuncSetAttribute(gemm2_bf16_tn_kernel<EPI_F32_RESIDUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM2_SMEM_BYTES));
  CK(cudaFuncSetAttribute(gemm2_bf16_tn_kernel<EPI_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM2_SMEM_BYTES));
  *out = e;
  return LEAF_OK;
}

static void free_workspace(leaf_engine* e) {
  cudaFree(e->x); cudaFree(e->h); cudaFree(e->big); cudaFree(e->pooled);
  cudaFree(e->cu); cudaFree(e->eos_row); cudaFree(e->total_rows); cudaFree(e->pfx); cudaFree(e->own_len); cudaFree(e->meta);
  e->x = nullptr; e->h = nullptr; e->big = nullptr; e->pooled = nullptr;
  e->cu = e->eos_row = e->total_rows = e->pfx = e->own_len = nullptr;
  e->meta = nullptr;
  e->max_seqs = 0; e->rows_cap = 0;
  e->tmaps.clear();
}

static void free_weights(leaf_engine* e) {
  for (auto& l : e->lw) {
    cudaFree(l.qkv_w); cudaFree(l.out_w); cudaFree(l.fc1_w); cudaFree(l.fc2_w); cudaFree(l.qkv_b_own);
  }
  e->lw.clear();
  cudaFree(e->proj_w);
  e->proj_w = nullptr;
  e->bound = false;
  e->tmaps.clear();
}

extern "C" int leaf_destroy(leaf_handle_t e) {
  if (!e) return LEAF_OK;
  cudaDeviceSynchronize();
  free_workspace(e);
  free_weights(e);
  for (void* p : e->table_allocs) cudaFree(p);
  for (auto& pr : e->gemm_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  for (auto ev : e->event_pool) cudaEventDestroy(ev);
  delete e;
  return LEAF_OK;
}

template <typename Tp>
static int upload(leaf_engine* e, const Tp* host, size_t count, const Tp** dev_out) {
  void* d = nullptr;
  CK(cudaMalloc(&d, count * sizeof(Tp)));
  e->table_allocs.push_back(d);
  CK(cudaMemcpy(d, host, count * sizeof(Tp), cudaMemcpyHostToDevice));
  *dev_out = static_cast<const Tp*>(d);
  return LEAF_OK;
}

extern "C" int leaf_load_bpe(leaf_handle_t e, const uint32_t* merge_pairs_host, int32_t n_merges) {
  if (!e || !merge_pairs_host) return fail(LEAF_ERR_INVALID, "null argument");
  if (n_merges != LEAF_N_MERGES) return fail(LEAF_ERR_INVALID, "CLIP's BPE has %d merges, got %d", LEAF_N_MERGES, n_merges);
  if (e->bpe_loaded) return LEAF_OK;
  std::vector<uint64_t> tab = k1_build_merge_table(merge_pairs_host, n_merges);
  K1Tables T{};
  int rc;
  if ((rc = upload(e, k1host::k1_host_byte_id, 256, &T.byte_id))) return rc;
  if ((rc = upload(e, k1host::k1_host_class, 256, &T.cls))) return rc;
  if ((rc = upload(e, k1host::k1_host_ws, 256, &T.ws))) return rc;
  if ((rc = upload(e, k1host::k1_host_lower, 256, &T.lower))) return rc;
  if ((rc = upload(e, k1host::k1_host_numref, 256, &T.numref))) return rc;
  if ((rc = upload(e, k1host::k1_host_ent_off, K1_N_ENTITIES, &T.ent_off))) return rc;
  if ((rc = upload(e, k1host::k1_host_ent_len, K1_N_ENTITIES, &T.ent_len))) return rc;
  if ((rc = upload(e, k1host::k1_host_ent_val, K1_N_ENTITIES, &T.ent_val))) return rc;
  if ((rc = upload(e, k1host::k1_host_ent_blob, K1_ENTITY_BLOB_BYTES, &T.ent_blob))) return rc;
  if ((rc = upload(e, tab.data(), tab.size(), &T.merge_tab))) return rc;
  T.n_ent = K1_N_ENTITIES;
  T.merge_bits = K1_MERGE_BITS;
  e->tables = T;
  e->bpe_loaded = true;
  return LEAF_OK;
}

static int cast_to(leaf_engine* e, const float* src, __nv_bfloat16* dst, size_t n, cudaStream_t st) {
  if (n % 4 != 0) return fail(LEAF_ERR_INVALID, "weight size not a multiple of 4");
  size_t blocks = (n / 4 + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  cast_bf16_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(src, dst, n);
  e->launches++;
  CK(cudaGetLastError());
  return LEAF_OK;
}

extern "C" int leaf_refresh_weights(leaf_handle_t e, void* stream) {
  if (!e || !e->bound) return fail(LEAF_ERR_STATE, "weights not bound");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t W = e->cfg.width, E = e->cfg.embed_dim;
  int rc;
  for (int l = 0; l < e->cfg.layers; ++l) {
    const leaf_layer_ptrs_t& p = e->layer_ptrs[l];
    LayerW& w = e->lw[l];
    if (p.in_proj_w) {
      if ((rc = cast_to(e, p.in_proj_w, w.qkv_w, 3 * W * W, st))) return rc;
      w.qkv_b = p.in_proj_b;
    } else {
      if ((rc = cast_to(e, p.q_w, w.qkv_w, W * W, st))) return rc;
      if ((rc = cast_to(e, p.k_w, w.qkv_w + W * W, W * W, st))) return rc;
      if ((rc = cast_to(e, p.v_w, w.qkv_w + 2 * W * W, W * W, st))) return rc;
      copy_f32_kernel<<<8, 256, 0, st>>>(p.q_b, w.qkv_b_own, W);
      copy_f32_kernel<<<8, 256, 0, st>>>(p.k_b, w.qkv_b_own + W, W);
      copy_f32_kernel<<<8, 256, 0, st>>>(p.v_b, w.qkv_b_own + 2 * W, W);
      e->launches += 3;
      w.qkv_b = w.qkv_b_own;
    }
    if ((rc = cast_to(e, p.out_w, w.out_w, W * W, st))) return rc;
    if ((rc = cast_to(e, p.fc1_w, w.fc1_w, 4 * W * W, st))) return rc;
    if ((rc = cast_to(e, p.fc2_w, w.fc2_w, 4 * W * W, st))) return rc;
  }
  if (e->wp.projection_is_ew) {
    if ((rc = cast_to(e, e->wp.text_projection, e->proj_w, E * W, st))) return rc;
  } else {
    dim3 grid(static_cast<unsigned>((E + 31) / 32), static_cast<unsigned>((W + 31) / 32));
    cast_bf16_transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(e->wp.text_projection, e->proj_w, static_cast<int>(W), static_cast<int>(E));
    e->launches++;
  }
  CK(cudaGetLastError());
  return LEAF_OK;
}

extern "C" int leaf_bind_weights(leaf_handle_t e, const leaf_weight_ptrs_t* w, void* stream) {
  if (!e || !w || !w->layers) return fail(LEAF_ERR_INVALID, "null argument");
  if (!w->token_embedding || !w->positional_embedding || !w->lnf_w || !w->lnf_b || !w->text_projection)
    return fail(LEAF_ERR_INVALID, "missing tower parameter");
  const size_t W = e->cfg.width, E = e->cfg.embed_dim;
  for (int l = 0; l < e->cfg.layers; ++l) {
    const leaf_layer_ptrs_t& p = w->layers[l];
    const bool fused = p.in_proj_w && p.in_proj_b;
    const bool split = p.q_w && p.k_w && p.v_w && p.q_b && p.k_b && p.v_b;
    if (!(fused || split) || !p.ln1_w || !p.ln1_b || !p.out_w || !p.out_b || !p.ln2_w || !p.ln2_b || !p.fc1_w || !p.fc1_b ||
        !p.fc2_w || !p.fc2_b)
      return fail(LEAF_ERR_INVALID, "missing parameter in layer %d", l);
  }
  free_weights(e);
  e->wp = *w;
  e->layer_ptrs.assign(w->layers, w->layers + e->cfg.layers);
  e->wp.layers = e->layer_ptrs.data();
  e->lw.resize(e->cfg.layers);
  for (auto& l : e->lw) {
    CK(cudaMalloc(&l.qkv_w, 3 * W * W * 2));
    CK(cudaMalloc(&l.out_w, W * W * 2));
    CK(cudaMalloc(&l.fc1_w, 4 * W * W * 2));
    CK(cudaMalloc(&l.fc2_w, 4 * W * W * 2));
    CK(cudaMalloc(&l.qkv_b_own, 3 * W * 4));
  }
  CK(cudaMalloc(&e->proj_w, E * W * 2));
  e->bound = true;
  return leaf_refresh_weights(e, stream);
}

extern "C" int leaf_reserve(leaf_handle_t e, int32_t max_seqs) {
  if (!e || max_seqs <= 0) return fail(LEAF_ERR_INVALID, "max_seqs");
  if (max_seqs <= e->max_seqs) return LEAF_OK;
  cudaDeviceSynchronize();
  free_workspace(e);
  const size_t W = e->cfg.width;
  const size_t rows = ((static_cast<size_t>(max_seqs) * LEAF_CTX + 127) / 128) * 128;
  CK(cudaMalloc(&e->x, rows * W * 4));
  CK(cudaMalloc(&e->h, rows * W * 2));
  CK(cudaMalloc(&e->big, rows * 4 * W * 2));
  CK(cudaMalloc(&e->pooled, (static_cast<size_t>(max_seqs) + 128) * W * 2));
  CK(cudaMalloc(&e->cu, (static_cast<size_t>(max_seqs) + 1) * 4));
  CK(cudaMalloc(&e->eos_row, static_cast<size_t>(max_seqs) * 4));
  CK(cudaMalloc(&e->pfx, static_cast<size_t>(max_seqs) * 4));
  CK(cudaMalloc(&e->own_len, static_cast<size_t>(max_seqs) * 4));
  CK(cudaMalloc(&e->meta, static_cast<size_t>(max_seqs) * 16));
  CK(cudaMalloc(&e->total_rows, 4));
  CK(cudaMemset(e->total_rows, 0, 4));
  e->max_seqs = max_seqs;
  e->rows_cap = static_cast<long>(rows);
  return LEAF_OK;
}

extern "C" int leaf_expand_tokenize(leaf_handle_t e, const uint8_t* caps, const int32_t* cap_off, int32_t B, int32_t n,
                                    const int32_t* pos, const int32_t* chr, const int32_t* sel, const uint8_t* valid,
                                    int32_t* tok_out, int32_t* len_out, int32_t* base_out, int32_t* status_out, void* stream) {
  if (!e || !caps || !cap_off || !tok_out || !len_out) return fail(LEAF_ERR_INVALID, "null argument");
  if (!e->bpe_loaded) return fail(LEAF_ERR_STATE, "leaf_load_bpe has not been called");
  if (B <= 0 || n < 0) return fail(LEAF_ERR_INVALID, "B=%d n=%d", B, n);
  if (n > 0 && (!pos || !chr)) return fail(LEAF_ERR_INVALID, "pos/chr required when n > 0");
  K1Args a{caps, cap_off, B, n, pos, chr, sel, valid, tok_out, len_out, base_out, status_out};
  const long R = static_cast<long>(B) * (n > 0 ? n : 1) + (n > 0 ? B : 0);
  const int grid = static_cast<int>((R + K1_WARPS_PER_CTA - 1) / K1_WARPS_PER_CTA);
  k1_expand_tokenize_kernel<<<grid, K1_WARPS_PER_CTA * 32, 0, static_cast<cudaStream_t>(stream)>>>(e->tables, a);
  e->launches++;
  CK(cudaGetLastError());
  return LEAF_OK;
}

static int launch_layernorm(leaf_engine* e, const float* x, const int* rows_dev, int rows_max, const int* gather,
                            const float* g, const float* b, __nv_bfloat16* y, cudaStream_t st) {
  const int W = e->cfg.width;
  const int vpl = W / 128;
  long warps = rows_max;
  int blocks = static_cast<int>((warps + 7) / 8);
  const int cap = e->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
#define LN_CASE(V) case V: layernorm_bf16_kernel<V><<<blocks, 256, 0, st>>>(x, rows_dev, rows_max, gather, W, g, b, e->cfg.ln_eps, y); break;
  switch (vpl) {
    LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8)
    LN_CASE(9) LN_CASE(10) LN_CASE(11) LN_CASE(12) LN_CASE(13) LN_CASE(14) LN_CASE(15) LN_CASE(16)
    default: return fail(LEAF_ERR_INVALID, "unsupported width %d", W);
  }
#undef LN_CASE
  e->launches++;
  CK(cudaGetLastError());
  return LEAF_OK;
}

extern "C" int leaf_encode(leaf_handle_t e, const int32_t* tok, const int32_t* len, const int32_t* base, int32_t N,
                           int32_t normalize, float* feat_out, void* stream) {
  if (!e || !tok || !len || !feat_out) return fail(LEAF_ERR_INVALID, "null argument");
  if (!e->bound) return fail(LEAF_ERR_STATE, "weights not bound");
  if (N <= 0) return fail(LEAF_ERR_INVALID, "N=%d", N);
  if (N > e->max_seqs) return fail(LEAF_ERR_STATE, "workspace reserved for %d rows, need %d (leaf_reserve)", e->max_seqs, N);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int W = e->cfg.width, E = e->cfg.embed_dim, H = e->cfg.heads;
  const int rows_max = static_cast<int>(static_cast<long>(N) * LEAF_CTX);
  int rc;
  prefix_kernel<<<(N + 7) / 8, 256, 0, st>>>(tok, len, base, N, e->pfx, e->own_len);
  scan_lengths_kernel<<<1, 1024, 0, st>>>(e->own_len, N, e->cu, e->total_rows);
  meta_kernel<<<(N + 255) / 256, 256, 0, st>>>(e->cu, e->pfx, base, N, e->meta, e->eos_row);
  embed_kernel<<<N, 256, 0, st>>>(tok, e->meta, N, W, e->wp.token_embedding, e->wp.positional_embedding, e->x);
  e->launches += 4;
  CK(cudaGetLastError());
  for (int l = 0; l < e->cfg.layers; ++l) {
    const leaf_layer_ptrs_t& p = e->layer_ptrs[l];
    const LayerW& w = e->lw[l];
    if ((rc = launch_layernorm(e, e->x, e->total_rows, rows_max, nullptr, p.ln1_w, p.ln1_b, e->h, st))) return rc;
    if ((rc = launch_gemm(e, e->h, e->rows_cap, w.qkv_w, w.qkv_b, e->big, 3 * W, rows_max, 3 * W, W, EPI_BF16, 0, e->total_rows, st))) return rc;
    attention_kernel<<<(N * H + ATT_WARPS - 1) / ATT_WARPS, ATT_WARPS * 32, 0, st>>>(e->big, e->meta, N, H, W, e->h);
    e->launches++;
    CK(cudaGetLastError());
    if ((rc = launch_gemm(e, e->h, e->rows_cap, w.out_w, p.out_b, e->x, W, rows_max, W, W, EPI_F32_RESIDUAL, 0, e->total_rows, st))) return rc;
    if ((rc = launch_layernorm(e, e->x, e->total_rows, rows_max, nullptr, p.ln2_w, p.ln2_b, e->h, st))) return rc;
    if ((rc = launch_gemm(e, e->h, e->rows_cap, w.fc1_w, p.fc1_b, e->big, 4 * W, rows_max, 4 * W, W, EPI_BF16_ACT, e->cfg.activation, e->total_rows, st))) return rc;
    if ((rc = launch_gemm(e, e->big, e->rows_cap, w.fc2_w, p.fc2_b, e->x, W, rows_max, W, 4 * W, EPI_F32_RESIDUAL, 0, e->total_rows, st))) return rc;
  }
  if ((rc = launch_layernorm(e, e->x, nullptr, N, e->eos_row, e->wp.lnf_w, e->wp.lnf_b, e->pooled, st))) return rc;
  if ((rc = launch_gemm(e, e->pooled, e->max_seqs + 128, e->proj_w, nullptr, feat_out, E, N, E, W, EPI_F32, 0, nullptr, st))) return rc;
  if (normalize) {
    l2_normalize_kernel<<<(N + 7) / 8, 256, 0, st>>>(feat_out, N, E);
    e->launches++;
  }
  CK(cudaGetLastError());
  return LEAF_OK;
}

extern "C" int leaf_score(leaf_handle_t e, const float* feat, const float* anchor, int32_t B, int32_t n, int32_t objective,
                          float* loss_out, int32_t* best_out, float* best_feat_out, void* stream) {
  if (!e || !feat || !anchor || !best_out) return fail(LEAF_ERR_INVALID, "null argument");
  if (B <= 0 || n <= 0 || objective < 0 || objective > 3) return fail(LEAF_ERR_INVALID, "B=%d n=%d objective=%d", B, n, objective);
  if (e->cfg.embed_dim % 4 != 0) return fail(LEAF_ERR_INVALID, "embed_dim");
  if (static_cast<size_t>(n) * 4 > 48 * 1024) return fail(LEAF_ERR_INVALID, "n=%d too large", n);
  score_argmax_kernel<<<B, 256, static_cast<size_t>(n) * 4, static_cast<cudaStream_t>(stream)>>>(
      feat, anchor, n, e->cfg.embed_dim, objective, loss_out, best_out, best_feat_out);
  e->launches++;
  CK(cudaGetLastError());
  return LEAF_OK;
}

extern "C" int leaf_gemm_bf16(leaf_handle_t e, const void* A, const void* Bt, const float* bias, void* C, int32_t M, int32_t N,
                              int32_t K, int32_t epilogue, int32_t act, const int32_t* m_dev, void* stream) {
  if (!e || !A || !Bt || !C) return fail(LEAF_ERR_INVALID, "null argument");
  return launch_gemm(e, static_cast<const __nv_bfloat16*>(A), M, static_cast<const __nv_bfloat16*>(Bt), bias, C, N, M, N, K,
                     epilogue, act, m_dev, static_cast<cudaStream_t>(stream));
}

extern "C" int leaf_test_layernorm(leaf_handle_t e, const float* x, int32_t rows, const float* gamma, const float* beta, void* y,
                                   void* stream) {
  if (!e || !x || !gamma || !beta || !y || rows <= 0) return fail(LEAF_ERR_INVALID, "bad argument");
  return launch_layernorm(e, x, nullptr, rows, nullptr, gamma, beta, static_cast<__nv_bfloat16*>(y), static_cast<cudaStream_t>(stream));
}

extern "C" int leaf_test_attention(leaf_handle_t e, const void* qkv, const int32_t* meta, int32_t N, void* out, void* stream) {
  if (!e || !qkv || !meta || !out || N <= 0) return fail(LEAF_ERR_INVALID, "bad argument");
  const int H = e->cfg.heads;
  attention_kernel<<<(N * H + ATT_WARPS - 1) / ATT_WARPS, ATT_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const int4*>(meta), N, H, e->cfg.width,
      static_cast<__nv_bfloat16*>(out));
  e->launches++;
  CK(cudaGetLastError());
  return LEAF_OK;
}

extern "C" int64_t leaf_launch_count(leaf_handle_t e, int32_t reset) {
  if (!e) return 0;
  const int64_t v = e->launches;
  if (reset) e->launches = 0;
  return v;
}

extern "C" int64_t leaf_last_rows(leaf_handle_t e) {
  if (!e || !e->total_rows) return 0;
  int v = 0;
  if (cudaMemcpy(&v, e->total_rows, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return v;
}

extern "C" int leaf_set_timing(leaf_handle_t e, int32_t on) {
  if (!e) return fail(LEAF_ERR_INVALID, "null handle");
  e->timing = on != 0;
  return LEAF_OK;
}

extern "C" double leaf_timing_ms(leaf_handle_t e, int32_t which, int32_t* launches) {
  if (!e || which != 0) return 0.0;
  double total = 0.0;
  int cnt = 0;
  for (auto& pr : e->gemm_events) {
    cudaEventSynchronize(pr.second);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) { total += ms; ++cnt; }
    e->event_pool.push_back(pr.first);
    e->event_pool.push_back(pr.second);
  }
  e->gemm_events.clear();
  if (launches) *launches = cnt;
  return total;
}
