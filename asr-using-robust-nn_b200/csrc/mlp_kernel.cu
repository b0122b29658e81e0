// Classifier forward pass of the accuracy-vs-SNR sweep: the reference's get_model
// (VDR/train_constraints.py:63-88: 880 -> 1024 -> 512 -> 256 -> 128 -> 64 (ReLU, BatchNormalization) -> 10 softmax) at inference,
// `model.predict` + argmax of VDR/attacks.py:409-414, as ONE launch.  BatchNormalization is folded into the following Dense layer
// by the host (asr_b200/mlp.py), so a layer is  h' = act(h W + b)  with W stored like Keras stores a Dense kernel: [in][out].
//
// mlp_forward_kernel: a CTA takes up to 32 rows through every layer (the host picks 8..32 so that every SM has a tile in each wave).  Activations live in ONE shared-memory buffer, k-major
// ([feature][32 rows]: the row values of one input feature are eight 16-byte broadcast loads); a layer's outputs stay in
// registers until every thread has finished reading its input, then overwrite it; the weights
// (6.4 MB for the reference's sizes) stay in L2 and are streamed once per CTA with loads that are coalesced along the output
// feature.  Thread t owns the output features t, t + 256, ... (at most kMlpMaxCols of them) for all 32 rows: up to 128 float32
// accumulators, the weight rows of the NEXT four k in flight (registers) while the current four are multiplied, one FMA per (row, feature, k) in ascending k - plain float32 like model.predict, no tensor cores: the sweep is
// bound by the MFCC launches (0.7 ms per 8192 clips and SNR), the 26 GFLOP of a forward pass over 8192 rows are 0.4 ms at the
// FP32 peak, and float32 accumulation keeps the decisions of the reference's float32 network.  The last layer's logits go through
// a max-subtracted softmax (expf) and an argmax per row.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include "common.cuh"

namespace asr {

constexpr int kMlpRowsMax = 32;       // rows per CTA: 8..32 in steps of 4 (shared memory for the widest activation, one tile per SM and wave)
constexpr int kMlpThreads = 256;
constexpr int kMlpMaxCols = 4;        // output features per thread: layers up to 1024 wide
constexpr int kMlpMaxLayers = 8;
constexpr int kMlpMaxWidth = kMlpThreads * kMlpMaxCols;

struct MlpParams {
  const float* x;                     // [n_rows][dims[0]] row-major
  long long n_rows;
  long long ld_x;
  int n_layers;
  int dims[kMlpMaxLayers + 1];
  const float* w[kMlpMaxLayers];      // [dims[l]][dims[l+1]]
  const float* b[kMlpMaxLayers];      // [dims[l+1]]
  float* logits;                      // [n_rows][dims[n_layers]] or null
  float* probs;                       // [n_rows][dims[n_layers]] or null
  int* argmax;                        // [n_rows] or null
  int buf_a, buf_b;                   // (unused: one activation buffer of max width x rows)
};

template <int NC, int kMlpRows>
__device__ __forceinline__ void mlp_layer(const float* in, float* out, const float* __restrict__ W,
                                          const float* __restrict__ bias, const int K, const int N, const bool relu) {
  const int t = threadIdx.x;
  float acc[NC][kMlpRows];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = t + c * kMlpThreads;
    const float b0 = col < N ? __ldg(bias + col) : 0.0f;
#pragma unroll
    for (int r = 0; r < kMlpRows; ++r) acc[c][r] = b0;
  }
  const float4* in4 = reinterpret_cast<const float4*>(in);
  // weight rows k..k+3 of this thread's features: loaded one group ahead of their use (L2 latency behind 256 FMAs)
  auto load_w = [&](float (&w)[4][NC], const int k0) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int col = t + c * kMlpThreads;
        w[u][c] = (col < N && k0 + u < K) ? __ldg(W + static_cast<long long>(k0 + u) * N + col) : 0.0f;
      }
  };
  auto fma_group = [&](const float (&w)[4][NC], const int k0) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (k0 + u < K) {
        float a[kMlpRows];
#pragma unroll
        for (int q = 0; q < kMlpRows / 4; ++q) {
          const float4 v = in4[(k0 + u) * (kMlpRows / 4) + q];        // broadcast: every thread reads the same 16 bytes
          a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int r = 0; r < kMlpRows; ++r) acc[c][r] = fmaf(a[r], w[u][c], acc[c][r]);
      }
    }
  };
  float w0[4][NC], w1[4][NC];
  load_w(w0, 0);
  for (int k = 0; k < K; k += 8) {                     // ascending k: the summation order of a plain float32 loop
    load_w(w1, k + 4);
    fma_group(w0, k);
    load_w(w0, k + 8);
    fma_group(w1, k + 4);
  }
  __syncthreads();                                     // every thread has read the layer's input: the outputs overwrite it
  float4* out4 = reinterpret_cast<float4*>(out);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = t + c * kMlpThreads;
    if (col < N) {
#pragma unroll
      for (int q = 0; q < kMlpRows / 4; ++q) {
        float4 v = make_float4(acc[c][4 * q], acc[c][4 * q + 1], acc[c][4 * q + 2], acc[c][4 * q + 3]);
        if (relu) { v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fmaxf(v.z, 0.0f); v.w = fmaxf(v.w, 0.0f); }
        out4[col * (kMlpRows / 4) + q] = v;
      }
    }
  }
}

template <int kMlpRows>
__global__ void __launch_bounds__(kMlpThreads) mlp_forward_kernel(const __grid_constant__ MlpParams mp) {
  extern __shared__ __align__(16) float mlp_smem[];
  float* const buf = mlp_smem;                         // ONE activation buffer: a layer's outputs stay in registers until its input is dead
  const long long r0 = static_cast<long long>(blockIdx.x) * kMlpRows;
  const int t = threadIdx.x;
  // ---- input rows -> k-major tile: a warp takes 32 features at a time, lane <-> feature: 16 coalesced row reads, then the
  //      lane's 16 row values as four 16-byte stores (rows past the end read as zero) ----
  {
    const int K = mp.dims[0], lane = t & 31, warp = t >> 5;
    for (int k0 = 32 * warp; k0 < K; k0 += 32 * (kMlpThreads / 32)) {
      const int k = k0 + lane;
      float v[kMlpRows];
#pragma unroll
      for (int r = 0; r < kMlpRows; ++r) {
        const long long row = r0 + r;
        v[r] = (k < K && row < mp.n_rows) ? __ldg(mp.x + row * mp.ld_x + k) : 0.0f;
      }
      if (k < K) {
        float4* dst = reinterpret_cast<float4*>(buf + k * kMlpRows);
#pragma unroll
        for (int q = 0; q < kMlpRows / 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    }
  }
  __syncthreads();
  for (int l = 0; l < mp.n_layers; ++l) {
    const int K = mp.dims[l], N = mp.dims[l + 1];
    const bool relu = l + 1 < mp.n_layers;
    const int nc = (N + kMlpThreads - 1) / kMlpThreads;
    const float* in = buf;
    float* out = buf;
    switch (nc) {
      case 1: mlp_layer<1, kMlpRows>(in, out, mp.w[l], mp.b[l], K, N, relu); break;
      case 2: mlp_layer<2, kMlpRows>(in, out, mp.w[l], mp.b[l], K, N, relu); break;
      case 3: mlp_layer<3, kMlpRows>(in, out, mp.w[l], mp.b[l], K, N, relu); break;
      default: mlp_layer<4, kMlpRows>(in, out, mp.w[l], mp.b[l], K, N, relu); break;
    }
    __syncthreads();
  }
  // ---- logits [class][32 rows] -> softmax / argmax per row (one thread per row) ----
  const int C = mp.dims[mp.n_layers];
  const float* lg = buf;
  if (t < kMlpRows && r0 + t < mp.n_rows) {
    const long long row = r0 + t;
    float mx = lg[t];
    int am = 0;
    for (int c = 1; c < C; ++c) {
      const float v = lg[c * kMlpRows + t];
      if (v > mx) { mx = v; am = c; }                   // first maximum, like np.argmax
    }
    if (mp.argmax) mp.argmax[row] = am;
    if (mp.logits)
      for (int c = 0; c < C; ++c) mp.logits[row * C + c] = lg[c * kMlpRows + t];
    if (mp.probs) {
      float s = 0.0f;
      for (int c = 0; c < C; ++c) s += expf(lg[c * kMlpRows + t] - mx);
      const float inv = 1.0f / s;
      for (int c = 0; c < C; ++c) mp.probs[row * C + c] = expf(lg[c * kMlpRows + t] - mx) * inv;
    }
  }
}

}  // namespace asr

using namespace asr;

// h_0 = x ; h_{l+1} = act(h_l W_l + b_l), act = ReLU except after the last layer ; probs = softmax(h_L), argmax = argmax(h_L).
// VDR/attacks.py:409-414 (model.predict + np.argmax) for the network of VDR/train_constraints.py:63-88 with BatchNormalization folded.
extern "C" int asr_mlp_forward(const float* x_dev, int64_t n_rows, int64_t ld_x, int32_t n_layers, const int32_t* dims_host,
                               const float* const* weights_dev, const float* const* biases_dev, float* logits_dev,
                               float* probs_dev, int32_t* argmax_dev, void* stream) {
  auto bad = [&](const char* m) { set_error(std::string("asr_mlp_forward: ") + m); return ASR_ERR_INVALID; };
  if (n_rows < 0) return bad("negative n_rows");
  if (n_rows == 0) return ASR_OK;
  if (!x_dev || !dims_host || !weights_dev || !biases_dev) return bad("null pointer");
  if (n_layers < 1 || n_layers > kMlpMaxLayers) return bad("1..8 layers");
  MlpParams mp;
  std::memset(&mp, 0, sizeof(mp));
  mp.x = x_dev; mp.n_rows = n_rows; mp.n_layers = n_layers;
  int wa = 0, wb = 0;                                    // widest activation held by buffer 0 (even layers' inputs) / buffer 1
  for (int l = 0; l <= n_layers; ++l) {
    const int d = dims_host[l];
    if (d < 1 || (l > 0 && d > kMlpMaxWidth)) return bad("layer widths must be 1..1024 (the input width is bounded by shared memory only)");
    mp.dims[l] = d;
    if (l % 2 == 0) wa = std::max(wa, d); else wb = std::max(wb, d);
    if (l < n_layers) {
      if (!weights_dev[l] || !biases_dev[l]) return bad("null weight / bias pointer");
      mp.w[l] = weights_dev[l]; mp.b[l] = biases_dev[l];
    }
  }
  if (ld_x < dims_host[0]) return bad("ld_x smaller than the input width");
  mp.ld_x = ld_x;
  mp.logits = logits_dev; mp.probs = probs_dev; mp.argmax = argmax_dev;
  const int widest = std::max(wa, wb);
  // rows per CTA: as many as fit (fewer weight passes over L2), but no more than it takes to give every SM a tile in each wave
  int dev = 0, sms = 148;
  ASR_CUDA_TRY(cudaGetDevice(&dev));
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) { cudaGetLastError(); sms = 148; }
  int cap = kMlpRowsMax;
  while (cap > 8 && static_cast<long long>(widest) * cap * 4 > kMaxSmemBytes) cap -= 4;
  if (static_cast<long long>(widest) * cap * 4 > kMaxSmemBytes) return bad("the widest activation (8 rows of it) does not fit one SM's shared memory");
  const long long waves = (n_rows + static_cast<long long>(sms) * cap - 1) / (static_cast<long long>(sms) * cap);
  long long want = (n_rows + sms * waves - 1) / (sms * waves);
  int rows = static_cast<int>(std::min<long long>(cap, std::max<long long>(8, (want + 3) / 4 * 4)));
  const int smem_bytes = static_cast<int>(static_cast<long long>(widest) * rows * 4);
  const long long tiles = (n_rows + rows - 1) / rows;
  if (tiles > 0x7fffffffLL) return bad("too many rows");
  const cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define ASR_MLP_LAUNCH(R)                                                                             \
  case R: {                                                                                           \
    static int granted[kMaxDevices] = {0};                                                            \
    ASR_CUDA_TRY(ensure_dyn_smem(mlp_forward_kernel<R>, smem_bytes, 48 * 1024, granted));             \
    mlp_forward_kernel<R><<<static_cast<unsigned>(tiles), kMlpThreads, smem_bytes, st>>>(mp);          \
  } break;
  switch (rows) {
    ASR_MLP_LAUNCH(8) ASR_MLP_LAUNCH(12) ASR_MLP_LAUNCH(16) ASR_MLP_LAUNCH(20) ASR_MLP_LAUNCH(24) ASR_MLP_LAUNCH(28)
    default: {
      static int granted[kMaxDevices] = {0};
      ASR_CUDA_TRY(ensure_dyn_smem(mlp_forward_kernel<32>, smem_bytes, 48 * 1024, granted));
      mlp_forward_kernel<32><<<static_cast<unsigned>(tiles), kMlpThreads, smem_bytes, st>>>(mp);
    } break;
  }
#undef ASR_MLP_LAUNCH
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}
