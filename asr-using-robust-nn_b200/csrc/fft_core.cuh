// Register FFT building blocks shared by the MFCC kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>

namespace asr {

// ------------------------------------------------------------------------------------------------
// compile-time twiddles exp(-2*pi*i*j/32), j = 0..15
__host__ __device__ constexpr float cos32(int j) {
  switch (j) {
    case 0: return 1.0f;
    case 1: return 0.98078528040323043f;
    case 2: return 0.92387953251128674f;
    case 3: return 0.83146961230254524f;
    case 4: return 0.70710678118654757f;
    case 5: return 0.55557023301960229f;
    case 6: return 0.38268343236508984f;
    case 7: return 0.19509032201612833f;
    case 8: return 0.0f;
    case 9: return -0.19509032201612819f;
    case 10: return -0.38268343236508973f;
    case 11: return -0.55557023301960196f;
    case 12: return -0.70710678118654746f;
    case 13: return -0.83146961230254535f;
    case 14: return -0.92387953251128674f;
    default: return -0.98078528040323043f;
  }
}
__host__ __device__ constexpr float sin32(int j) {
  switch (j) {
    case 0: return 0.0f;
    case 1: return 0.19509032201612825f;
    case 2: return 0.38268343236508978f;
    case 3: return 0.55557023301960218f;
    case 4: return 0.70710678118654746f;
    case 5: return 0.83146961230254524f;
    case 6: return 0.92387953251128674f;
    case 7: return 0.98078528040323043f;
    case 8: return 1.0f;
    case 9: return 0.98078528040323043f;
    case 10: return 0.92387953251128674f;
    case 11: return 0.83146961230254546f;
    case 12: return 0.70710678118654757f;
    case 13: return 0.55557023301960218f;
    case 14: return 0.38268343236508989f;
    default: return 0.19509032201612861f;
  }
}

template <int P>
__host__ __device__ constexpr int brev(int i) {
  int r = 0;
  for (int b = 1; b < P; b <<= 1) {
    r = (r << 1) | (i & 1);
    i >>= 1;
  }
  return r;
}

// In-register radix-2 decimation-in-time DFT of P complex points (P <= 32).
// Input in bit-reversed order, output in natural order.  Non-trivial butterflies use the
// 6-FMA form  a' = a + w*b ; b' = 2a - a'.
template <int P, int M0 = 2>   // M0 = 4: the first stage (m = 2) has already been applied
__device__ __forceinline__ void dft_dit_from(float (&re)[P], float (&im)[P]) {
#pragma unroll
  for (int m = M0; m <= P; m *= 2) {
    const int h = m / 2;
#pragma unroll
    for (int g = 0; g < P; g += m) {
#pragma unroll
      for (int j = 0; j < h; ++j) {
        const int a = g + j, b = a + h;
        const int tw = j * (32 / m);
        const float ar = re[a], ai = im[a], br = re[b], bi = im[b];
        if (tw == 0) {
          re[a] = ar + br; im[a] = ai + bi;
          re[b] = ar - br; im[b] = ai - bi;
        } else if (tw == 8) {   // w = -i : w*b = (bi, -br)
          re[a] = ar + bi; im[a] = ai - br;
          re[b] = ar - bi; im[b] = ai + br;
        } else {
          const float wr = cos32(tw), wi = -sin32(tw);
          float nr = fmaf(wr, br, ar);
          nr = fmaf(-wi, bi, nr);
          float ni = fmaf(wr, bi, ai);
          ni = fmaf(wi, br, ni);
          re[a] = nr; im[a] = ni;
          re[b] = fmaf(2.0f, ar, -nr);
          im[b] = fmaf(2.0f, ai, -ni);
        }
      }
    }
  }
}

template <int P>
__device__ __forceinline__ void dft_dit(float (&re)[P], float (&im)[P]) { dft_dit_from<P, 2>(re, im); }

template <int NFFT> struct FftCfg;
template <> struct FftCfg<512>  { static constexpr int M = 256,  G = 16, P = 16; };
template <> struct FftCfg<1024> { static constexpr int M = 512,  G = 16, P = 32; };
template <> struct FftCfg<2048> { static constexpr int M = 1024, G = 32, P = 32; };

// Power spectrum of one real frame of NFFT samples via a complex FFT of M = NFFT/2 points held by
// a group of G lanes (P = M/G points per lane):   n = n1 + G*n2 ,  k = k2 + P*k1
//   pass 1 (lane n1): DFT_P over n2, times W_M^(n1*k2)            -> smem exchange
//   pass 2 (lane l ): DFT_G over n1 for k2 = l + G*q              -> Z[k]
//   unpack          : X[k], X[M-k] from Z[k], conj Z[M-k] (fetched from lane G-l by shuffle)
//                     -> |X|^2 into fbuf[0..M]
// fbuf is the exchange buffer, srow receives the power spectrum (they may be the same buffer).
// HALF_XB: the exchange goes through a buffer of G/2 rows in two rounds (lanes 0..G/2-1 write, everybody reads its
// first G/2 operands, then the upper lanes): half the shared memory for G/2 more (half-populated) store instructions.
template <int NFFT, bool ZERO_TAIL, bool HALF_XB = false>
__device__ __forceinline__ void frame_power_fft(float (&re)[FftCfg<NFFT>::P], float (&im)[FftCfg<NFFT>::P],
                                                const float* __restrict__ twp, const float2* __restrict__ twu,
                                                float* fbuf, float* srow, const int l) {
  constexpr int M = FftCfg<NFFT>::M, G = FftCfg<NFFT>::G, P = FftCfg<NFFT>::P;
  constexpr int Q = P / G;
  float2* xb = reinterpret_cast<float2*>(fbuf);
  dft_dit<P>(re, im);
  {
    const float4* tw4 = reinterpret_cast<const float4*>(twp + l * (2 * P + 4));
#pragma unroll
    for (int k2 = 0; k2 < P; k2 += 2) {
      const float4 t = tw4[k2 / 2];
      if (k2 != 0) {
        const float r = re[k2], i = im[k2];
        re[k2] = fmaf(r, t.x, -i * t.y);
        im[k2] = fmaf(r, t.y, i * t.x);
      }
      const float r = re[k2 + 1], i = im[k2 + 1];
      re[k2 + 1] = fmaf(r, t.z, -i * t.w);
      im[k2 + 1] = fmaf(r, t.w, i * t.z);
    }
  }
  float ur[Q][G], ui[Q][G];
  if constexpr (!HALF_XB) {
    __syncwarp();                                  // staged frames read their samples from the frame buffers
#pragma unroll
    for (int k2 = 0; k2 < P; ++k2) xb[l * (P + 1) + k2] = make_float2(re[k2], im[k2]);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < Q; ++q) {
#pragma unroll
      for (int n1 = 0; n1 < G; ++n1) {
        const float2 a = xb[n1 * (P + 1) + l + G * q];
        ur[q][brev<G>(n1)] = a.x;
        ui[q][brev<G>(n1)] = a.y;
      }
    }
  } else {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      __syncwarp();                                // round 0: earlier readers of the buffer; round 1: round 0's reads
      if ((l >= G / 2) == (half == 1)) {
        const int row = l - half * (G / 2);
#pragma unroll
        for (int k2 = 0; k2 < P; ++k2) xb[row * (P + 1) + k2] = make_float2(re[k2], im[k2]);
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < Q; ++q) {
#pragma unroll
        for (int r = 0; r < G / 2; ++r) {
          const float2 a = xb[r * (P + 1) + l + G * q];
          ur[q][brev<G>(half * (G / 2) + r)] = a.x;
          ui[q][brev<G>(half * (G / 2) + r)] = a.y;
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < Q; ++q) dft_dit<G>(ur[q], ui[q]);
  __syncwarp();                                    // exchange data consumed; fbuf becomes the spectrum
  const int partner = (G - l) & (G - 1);
#pragma unroll
  for (int q = 0; q < Q; ++q)
#pragma unroll
    for (int k1 = 0; k1 < G / 2; ++k1) {
      const int k = l + G * q + P * k1;
      // Z[M-k] lives in lane G-l at (Q-1-q, G-1-k1); lane 0 pairs with itself at (Q-q, G-1-k1) for q >= 1
      // and at (0, (G-k1) mod G) for q = 0.  Every lane offers what its requester needs.
      const int q0 = (q == 0) ? 0 : Q - q, j0 = (q == 0) ? ((G - k1) & (G - 1)) : G - 1 - k1;
      const float give_r = (l == 0) ? ur[q0][j0] : ur[Q - 1 - q][G - 1 - k1];
      const float give_i = (l == 0) ? ui[q0][j0] : ui[Q - 1 - q][G - 1 - k1];
      const float br = __shfl_sync(0xffffffffu, give_r, partner, G);
      const float bi = __shfl_sync(0xffffffffu, give_i, partner, G);
      const float2 w = twu[k];                 // (-sin(2 pi k/N)/2, -cos(2 pi k/N)/2)
      const float ar = ur[q][k1], ai = ui[q][k1];
      const float sr = ar + br, si = ai - bi;  // A + conj(B)
      const float dr = ar - br, di = ai + bi;  // A - conj(B)
      const float tr = fmaf(w.x, dr, -w.y * di);
      const float ti = fmaf(w.x, di, w.y * dr);
      const float xr = fmaf(0.5f, sr, tr), xi = fmaf(0.5f, si, ti);
      const float yr = fmaf(0.5f, sr, -tr), yi = fmaf(0.5f, si, -ti);
      srow[k] = fmaf(xr, xr, xi * xi);
      srow[M - k] = fmaf(yr, yr, yi * yi);
    }
  if (l == 0) srow[M / 2] = fmaf(ur[0][G / 2], ur[0][G / 2], ui[0][G / 2] * ui[0][G / 2]);
  if (ZERO_TAIL && l < 16) srow[M + 1 + l] = 0.0f;   // tail read by the (fixed 4-quad) mel tasks
}

}  // namespace asr
