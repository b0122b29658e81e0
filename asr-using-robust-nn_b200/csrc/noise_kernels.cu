// Additive-noise path for sm_100a (reference: VDR/attacks.py:73-86,145-183,222-245;
// SR/attacks.py:81-94,149-189,228-251).
//
//   clip_power_kernel  P = np.mean(sample**2) in float32, bit-exact with numpy's pairwise summation
//   snr_sigma_kernel   the float32 scalar chain of add_white_noise_with_snr
//   mix_*_kernel       float64(x) + s*z with two separately rounded float64 operations
//   randn_kernel       seeded Philox4x32-10 + Box-Muller standard-normal stream
#include "common.cuh"

namespace asr {

// ------------------------------------------------------------------------------------------------
// numpy pairwise_sum(a, n):
//   n < 8    : serial
//   n <= 128 : 8 strided accumulators, ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail serially
//   else     : n2 = n/2 - (n/2 % 8) ; pairwise(a, n2) + pairwise(a+n2, n-n2)
// The recursion tree only depends on n.  One WARP per clip walks it depth-first with an explicit
// (offset, length, depth) stack, takes the leaves (n <= 128) four at a time - one leaf per group of 8
// lanes, lane j of a group owning numpy's accumulator r[j] - and folds the leaf sums with a
// (value, depth) stack: two entries of equal depth are siblings and merge into their parent, which is
// exactly the order numpy adds them in.  No shared-memory staging, no CTA barriers.
constexpr int kPowWarpsWide = 12;  // CTA shapes of clip_power_kernel: one CTA of 12 warps per SM, or three of 4 (same warps per SM, finer grain)
constexpr int kPowWarpsNarrow = 4;
constexpr int kPowDefaultWarps = kPowWarpsNarrow;   // measured: 0.062 ms against 0.073 ms per 8192 one-second clips stand-alone, and the better partner of a step's tail kernels (profiles/r2_power_overlap_ab.txt)
constexpr int kPowStack = 40;     // > depth of the tree for any int32 length

struct PowScratch {               // per warp
  int s_off[kPowStack], s_len[kPowStack], s_dep[kPowStack];
  float v_val[kPowStack];
  int v_dep[kPowStack];
  int l_off[4], l_len[4], l_dep[4];
};

template <int DT>
__device__ __forceinline__ float load_sq(const void* __restrict__ audio, const long long i) {
  float x;
  if (DT == ASR_I16) x = static_cast<float>(__ldg(reinterpret_cast<const short*>(audio) + i)) * (1.0f / 32768.0f);
  else x = __ldg(reinterpret_cast<const float*>(audio) + i);
  return __fmul_rn(x, x);        // sample**2, exact float32 product
}

// One leaf (n <= 128) summed by a group of 8 lanes; the xor-shuffle tree reproduces
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) (float addition is commutative), lane 0 of the group adds the tail.
template <int DT>
__device__ __forceinline__ float leaf_sum8(const void* __restrict__ audio, const long long e0, const int n, const int j) {
  float r = 0.0f;
  if (n < 8) {
    if (j == 0)
      for (int i = 0; i < n; ++i) r = __fadd_rn(r, load_sq<DT>(audio, e0 + i));
  } else {
    const int n8 = n - (n % 8);
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (8 * i < n8) ? load_sq<DT>(audio, e0 + 8 * i + j) : 0.0f;
    r = v[0];
#pragma unroll
    for (int i = 1; i < 16; ++i)
      if (8 * i < n8) r = __fadd_rn(r, v[i]);
  }
  // all 32 lanes shuffle (groups with no leaf carry zeros); xor < 8 stays inside the group
  r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
  r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
  r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
  if (n >= 8 && j == 0) {
    const int n8 = n - (n % 8);
    for (int i = n8; i < n; ++i) r = __fadd_rn(r, load_sq<DT>(audio, e0 + i));
  }
  return r;
}

// Leaf table of a length: (offset, length, merges after the leaf) in left-to-right order.  Built once per CTA for
// the length of its first clip (a batch usually holds one length) and replayed by every warp whose clip has that
// length; other clips walk the tree themselves (clip_power_walk).
constexpr int kPowMaxLeaves = 640;       // covers lengths up to ~80 000 samples; longer clips take the walking path

struct PowTable {
  int n_leaves, length;
  int perfect;                           // 1: 2^k >= 32 leaves, all at the same depth, all of whole rows of 8 (the tree is a perfect binary tree)
  unsigned short off8[kPowMaxLeaves];    // leaf offset / 8 (every leaf starts on a multiple of 8)
  unsigned char len[kPowMaxLeaves];      // leaf length, 1..128 -> stored minus 1... (0 marks an empty clip)
  unsigned char merges[kPowMaxLeaves];
};

// Built by one warp, level by level: every node longer than 128 splits the way numpy does, the order of the nodes is
// kept (ballot + popcount give the output positions), until only leaves remain.  `buf` = 6 * kPowMaxLeaves ints.
__device__ void pow_table_build(PowTable& tb, const int L, int* __restrict__ buf, const int lane) {
  constexpr int cap = kPowMaxLeaves;
  int* A = buf;             // [off | len | depth] x cap
  int* B = buf + 3 * cap;
  bool ok = L > 0 && L <= 8 * 65535;
  int n = 1;
  if (lane == 0) { A[0] = 0; A[cap] = L; A[2 * cap] = 0; }
  __syncwarp();
  bool any = ok && L > 128;
  while (any) {
    int out = 0;
    bool more = false, overflow = false;
    for (int c = 0; c < n; c += 32) {
      const int i = c + lane;
      const bool valid = i < n;
      const int off = valid ? A[i] : 0, len = valid ? A[cap + i] : 0, d = valid ? A[2 * cap + i] : 0;
      const bool split = len > 128;
      const unsigned m = __ballot_sync(0xffffffffu, split);
      const int pos = out + lane + __popc(m & ((1u << lane) - 1u));
      if (valid) {
        if (pos + 1 >= cap) {
          overflow = true;
        } else if (split) {
          int n2 = len / 2;
          n2 -= n2 % 8;
          B[pos] = off;          B[cap + pos] = n2;           B[2 * cap + pos] = d + 1;
          B[pos + 1] = off + n2; B[cap + pos + 1] = len - n2; B[2 * cap + pos + 1] = d + 1;
          more = more || n2 > 128 || len - n2 > 128;
        } else {
          B[pos] = off; B[cap + pos] = len; B[2 * cap + pos] = d;
        }
      }
      out += min(32, n - c) + __popc(m);
    }
    if (__any_sync(0xffffffffu, overflow)) { ok = false; break; }
    any = __any_sync(0xffffffffu, more);
    n = out;
    int* t = A; A = B; B = t;
    __syncwarp();
  }
  if (ok && n <= cap) {
    for (int i = lane; i < n; i += 32) {
      tb.off8[i] = static_cast<unsigned short>(A[i] >> 3);
      tb.len[i] = static_cast<unsigned char>(A[cap + i] - 1);
    }
    // perfect tree: 2^k >= 32 leaves, all at the same depth, whole rows of 8 (checked by all lanes)
    bool perfect = n >= 32 && (n & (n - 1)) == 0;
    for (int i = lane; i < n; i += 32) perfect = perfect && A[2 * cap + i] == A[2 * cap] && (A[cap + i] & 7) == 0;
    perfect = __all_sync(0xffffffffu, perfect);
    if (perfect)                                       // a perfect tree merges like a binary counter: trailing ones of the leaf index
      for (int i = lane; i < n; i += 32) tb.merges[i] = static_cast<unsigned char>(__ffs(~i) - 1);
    __syncwarp();
    if (lane == 0) {
      if (!perfect) {
        // merges after each leaf: the depths on the value stack are strictly increasing -> a bit mask is the stack
        unsigned stack = 0;
        for (int i = 0; i < n; ++i) {
          int d = A[2 * cap + i], k = 0;
          while (stack & (1u << d)) { stack &= ~(1u << d); --d; ++k; }
          stack |= 1u << d;
          tb.merges[i] = static_cast<unsigned char>(k);
        }
      }
      tb.n_leaves = n;
      tb.length = L;
      tb.perfect = perfect ? 1 : 0;
    }
  } else if (lane == 0) {
    tb.n_leaves = 0;
    tb.length = -1;
    tb.perfect = 0;
  }
}

// generic path: one warp walks the tree of its clip (any length)
template <int DT>
__device__ void clip_power_walk(const void* __restrict__ audio, const long long base, const int L, PowScratch& sc,
                                const int lane, float* __restrict__ out) {
  const int g = lane >> 3, j = lane & 7;
  int sp = 0, vp = 0;                             // stack pointers (uniform over the warp)
  if (lane == 0) { sc.s_off[0] = 0; sc.s_len[0] = L; sc.s_dep[0] = 0; }
  sp = 1;
  __syncwarp();
  while (sp > 0) {
    // ---- lane 0 pops nodes until it holds up to 4 leaves (left-to-right order) ----
    int n_leaf = 0;
    if (lane == 0) {
      while (sp > 0 && n_leaf < 4) {
        --sp;
        const int off = sc.s_off[sp], n = sc.s_len[sp], d = sc.s_dep[sp];
        if (n > 128) {
          int n2 = n / 2;
          n2 -= n2 % 8;
          sc.s_off[sp] = off + n2; sc.s_len[sp] = n - n2; sc.s_dep[sp] = d + 1;   // right child, visited later
          sc.s_off[sp + 1] = off;  sc.s_len[sp + 1] = n2; sc.s_dep[sp + 1] = d + 1;
          sp += 2;
        } else {
          sc.l_off[n_leaf] = off; sc.l_len[n_leaf] = n; sc.l_dep[n_leaf] = d;
          ++n_leaf;
        }
      }
    }
    n_leaf = __shfl_sync(0xffffffffu, n_leaf, 0);
    sp = __shfl_sync(0xffffffffu, sp, 0);
    __syncwarp();
    float r = 0.0f;
    {
      const bool have = g < n_leaf;
      const int off = have ? sc.l_off[g] : 0, n = have ? sc.l_len[g] : 0;
      r = leaf_sum8<DT>(audio, base + off, n, j);
    }
    // ---- fold the leaf sums: equal depth on top of the value stack = siblings ----
    const float r0 = __shfl_sync(0xffffffffu, r, 0), r1 = __shfl_sync(0xffffffffu, r, 8);
    const float r2 = __shfl_sync(0xffffffffu, r, 16), r3 = __shfl_sync(0xffffffffu, r, 24);
    if (lane == 0) {
      const float rs[4] = {r0, r1, r2, r3};
      for (int q = 0; q < n_leaf; ++q) {
        float v = rs[q];
        int d = sc.l_dep[q];
        while (vp > 0 && sc.v_dep[vp - 1] == d) {
          --vp;
          v = __fadd_rn(sc.v_val[vp], v);         // left + right
          --d;
        }
        sc.v_val[vp] = v; sc.v_dep[vp] = d;
        ++vp;
      }
    }
    vp = __shfl_sync(0xffffffffu, vp, 0);
    __syncwarp();
  }
  // np.mean: float32 sum / count evaluated in float64, rounded to float32
  if (lane == 0) *out = static_cast<float>(static_cast<double>(sc.v_val[0]) / static_cast<double>(L));
}

// table path: leaves four at a time (one per group of 8 lanes), the loads of the next four are in flight while
// the current four are summed; lane 0 replays the merge counts.
template <int DT>
__device__ void clip_power_replay(const void* __restrict__ audio, const long long base, const int L, const PowTable& tb,
                                  PowScratch& sc, const int lane, float* __restrict__ out) {
  const int g = lane >> 3, j = lane & 7;
  const int nl = tb.n_leaves;
  float cur[16], nxt[16];
  auto fetch = [&](const int leaf, float (&v)[16]) {
    const bool have = leaf < nl;
    const int off = have ? 8 * tb.off8[leaf] : 0, n = have ? tb.len[leaf] + 1 : 0;
    const int n8 = n - (n % 8);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (8 * i < n8) ? load_sq<DT>(audio, base + off + 8 * i + j) : 0.0f;
  };
  int vp = 0;
  fetch(g, cur);
  for (int l0 = 0; l0 < nl; l0 += 4) {
    fetch(l0 + 4 + g, nxt);
    const int leaf = l0 + g;
    const bool have = leaf < nl;
    const int off = have ? 8 * tb.off8[leaf] : 0, n = have ? tb.len[leaf] + 1 : 0;
    const int n8 = n - (n % 8);
    float r = 0.0f;
    if (n >= 8) {
      r = cur[0];
#pragma unroll
      for (int i = 1; i < 16; ++i)
        if (8 * i < n8) r = __fadd_rn(r, cur[i]);
    }
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
    if (j == 0) {
      if (n < 8) { r = 0.0f; for (int i = 0; i < n; ++i) r = __fadd_rn(r, load_sq<DT>(audio, base + off + i)); }
      else for (int i = n8; i < n; ++i) r = __fadd_rn(r, load_sq<DT>(audio, base + off + i));
    }
    const float r0 = __shfl_sync(0xffffffffu, r, 0), r1 = __shfl_sync(0xffffffffu, r, 8);
    const float r2 = __shfl_sync(0xffffffffu, r, 16), r3 = __shfl_sync(0xffffffffu, r, 24);
    if (lane == 0) {
      const float rs[4] = {r0, r1, r2, r3};
      for (int q = 0; q < 4 && l0 + q < nl; ++q) {
        float v = rs[q];
        for (int k = tb.merges[l0 + q]; k > 0; --k) v = __fadd_rn(sc.v_val[--vp], v);   // left + right
        sc.v_val[vp++] = v;
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
  }
  __syncwarp();
  if (lane == 0) *out = static_cast<float>(static_cast<double>(sc.v_val[0]) / static_cast<double>(L));
}

// Vector variant of the table path for 16-byte aligned clips: a group of 8 lanes loads its leaf row by row with
// 16-byte loads (lane j: rows j and j+8 of 8 samples), squares go through a per-warp shared-memory tile
// ([row][8], pitch 12: conflict-free both ways), lane j then reads column j = numpy's accumulator r[j].
// Rows past the end of a short leaf hold +0.0; adding it is exact, so the row loop needs no predicates.
constexpr int kPowTileGroup = 200;        // floats per group tile (16 rows x pitch 12, + 8: groups 8 banks apart)

template <int DT> struct PowRaw;
template <> struct PowRaw<ASR_I16> { int4 a, b; };
template <> struct PowRaw<ASR_F32> { float4 a0, a1, b0, b1; };

template <int DT>
__device__ __forceinline__ void pow_rows_load(const void* __restrict__ audio, const long long e, const bool va,
                                              const bool vb, PowRaw<DT>& r) {
  if constexpr (DT == ASR_I16) {
    const int4* p = reinterpret_cast<const int4*>(reinterpret_cast<const short*>(audio) + e);
    r.a = va ? __ldg(p) : make_int4(0, 0, 0, 0);
    r.b = vb ? __ldg(p + 8) : make_int4(0, 0, 0, 0);
  } else {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(audio) + e);
    const float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    r.a0 = va ? __ldg(p) : z; r.a1 = va ? __ldg(p + 1) : z;
    r.b0 = vb ? __ldg(p + 16) : z; r.b1 = vb ? __ldg(p + 17) : z;
  }
}

// squares of one row of 8 samples -> two float4
template <int DT>
__device__ __forceinline__ void pow_row_squares(const PowRaw<DT>& r, const int which, float4& lo, float4& hi) {
  float v[8];
  if constexpr (DT == ASR_I16) {
    const int4 q = which ? r.b : r.a;
    const unsigned w[4] = {static_cast<unsigned>(q.x), static_cast<unsigned>(q.y), static_cast<unsigned>(q.z),
                           static_cast<unsigned>(q.w)};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned x = w[k] ^ 0x80008000u;    // exact int16 -> float: bits(2^23 + (s + 32768)) - (2^23 + 32768)
      const float s0 = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7410)) - 8421376.0f;
      const float s1 = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7432)) - 8421376.0f;
      // (s/32768)^2 rounded to float32 == round(s*s) * 2^-30 (scaling by a power of two commutes with rounding)
      v[2 * k] = __fmul_rn(__fmul_rn(s0, s0), 9.313225746154785e-10f);
      v[2 * k + 1] = __fmul_rn(__fmul_rn(s1, s1), 9.313225746154785e-10f);
    }
  } else {
    const float4 x0 = which ? r.b0 : r.a0, x1 = which ? r.b1 : r.a1;
    v[0] = __fmul_rn(x0.x, x0.x); v[1] = __fmul_rn(x0.y, x0.y); v[2] = __fmul_rn(x0.z, x0.z); v[3] = __fmul_rn(x0.w, x0.w);
    v[4] = __fmul_rn(x1.x, x1.x); v[5] = __fmul_rn(x1.y, x1.y); v[6] = __fmul_rn(x1.z, x1.z); v[7] = __fmul_rn(x1.w, x1.w);
  }
  lo = make_float4(v[0], v[1], v[2], v[3]);
  hi = make_float4(v[4], v[5], v[6], v[7]);
}

template <int DT>
__device__ void clip_power_replay_vec(const void* __restrict__ audio, const long long base, const int L,
                                      const PowTable& tb, PowScratch& sc, float* __restrict__ tile, const int lane,
                                      float* __restrict__ out) {
  const int g = lane >> 3, j = lane & 7;
  const int nl = tb.n_leaves;
  float* tg = tile + g * kPowTileGroup;
  PowRaw<DT> cur, nxt;
  auto issue = [&](const int leaf, PowRaw<DT>& r) {
    const bool have = leaf < nl;
    const int off = have ? 8 * tb.off8[leaf] : 0, n = have ? tb.len[leaf] + 1 : 0;
    const int rows = n >> 3;                       // full rows of 8
    pow_rows_load<DT>(audio, base + off + 8 * j, j < rows, j + 8 < rows, r);
  };
  int vp = 0;
  issue(g, cur);
  for (int l0 = 0; l0 < nl; l0 += 4) {
    issue(l0 + 4 + g, nxt);
    {
      float4 lo, hi;
      pow_row_squares<DT>(cur, 0, lo, hi);
      *reinterpret_cast<float4*>(tg + 12 * j) = lo;
      *reinterpret_cast<float4*>(tg + 12 * j + 4) = hi;
      pow_row_squares<DT>(cur, 1, lo, hi);
      *reinterpret_cast<float4*>(tg + 12 * (j + 8)) = lo;
      *reinterpret_cast<float4*>(tg + 12 * (j + 8) + 4) = hi;
    }
    __syncwarp();
    float r = tg[j];
#pragma unroll
    for (int i = 1; i < 16; ++i) r = __fadd_rn(r, tg[12 * i + j]);
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
    {
      const int leaf = l0 + g;
      const int n = leaf < nl ? tb.len[leaf] + 1 : 0;
      if ((n & 7) != 0 && j == 0) {                 // tail of the (last) leaf, serially like numpy; n < 8: the whole leaf
        const int off = 8 * tb.off8[leaf];
        if (n < 8) r = 0.0f;
        for (int i = n & ~7; i < n; ++i) r = __fadd_rn(r, load_sq<DT>(audio, base + off + i));
      }
    }
    const float r0 = __shfl_sync(0xffffffffu, r, 0), r1 = __shfl_sync(0xffffffffu, r, 8);
    const float r2 = __shfl_sync(0xffffffffu, r, 16), r3 = __shfl_sync(0xffffffffu, r, 24);
    if (lane == 0) {
      const float rs[4] = {r0, r1, r2, r3};
      for (int q = 0; q < 4 && l0 + q < nl; ++q) {
        float v = rs[q];
        for (int k = tb.merges[l0 + q]; k > 0; --k) v = __fadd_rn(sc.v_val[--vp], v);   // left + right
        sc.v_val[vp++] = v;
      }
    }
    cur = nxt;
    __syncwarp();                                   // the tile is rewritten next round
  }
  if (lane == 0) *out = static_cast<float>(static_cast<double>(sc.v_val[0]) / static_cast<double>(L));
}

// Perfect trees (e.g. 16 000 samples: 128 leaves of 120 / 128 at depth 7), int16, 16-byte aligned clips: lanes <-> LEAVES.
// A round takes 32 consecutive leaves: their samples (one contiguous range, <= 8 KB) are copied to shared memory with
// coalesced 16-byte loads (a 16-byte pad after every 16 rows spreads the lanes' rows over the banks), then every lane sums
// ITS leaf exactly as numpy does - eight accumulators r[j] += a[8i + j] over the rows i in order, then
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) - and the 32 leaf sums fold by xor-shuffles: neighbours are siblings at every level of
// a perfect tree (float addition is commutative, so both partners hold the parent).  Rounds fold through a small stack
// the same way.  The samples stay UNSCALED integers in float32: scaling every operand by 2^-30 commutes with every
// rounding of the sum (no underflow: a nonzero square is >= 1), so one exact multiply at the end replaces 16 000.
constexpr int kPowPerfTile = 2176;        // floats per tile: 512 rows of 16 bytes + 32 pads; two tiles per warp
__device__ __forceinline__ void pow_cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
template <int DT>
__device__ void clip_power_perfect(const void* __restrict__ audio, const long long base, const int L, const PowTable& tb,
                                   PowScratch& sc, float* __restrict__ tile, const int lane, float* __restrict__ out) {
  static_assert(DT == ASR_I16, "int16 only");
  const int nl = tb.n_leaves;
  const int4* src = reinterpret_cast<const int4*>(reinterpret_cast<const short*>(audio) + base);   // rows of 8 samples
  // rows of a round -> its tile, asynchronously (cp.async: no registers, the copy of round r+1 runs under the sums of round r)
  auto issue = [&](const int l0, char* dst) {
    const int q0 = tb.off8[l0];                                             // first row of the round
    const int nq = tb.off8[l0 + 31] + ((tb.len[l0 + 31] + 1) >> 3) - q0;    // rows of the round, <= 512
    for (int q = lane; q < nq; q += 32) pow_cp_async16(dst + 16 * (q + (q >> 4)), src + q0 + q);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int sp = 0;                                       // round stack (uniform over the warp)
  __syncwarp();                                     // the tiles are rewritten
  issue(0, reinterpret_cast<char*>(tile));
  for (int l0 = 0, round = 0; l0 < nl; l0 += 32, ++round) {
    const char* tb8 = reinterpret_cast<const char*>(tile + (round & 1) * kPowPerfTile);
    if (l0 + 32 < nl) {
      issue(l0 + 32, reinterpret_cast<char*>(tile + ((round + 1) & 1) * kPowPerfTile));
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    const int leaf = l0 + lane;
    const int qs = tb.off8[leaf] - tb.off8[l0], rows = (tb.len[leaf] + 1) >> 3;      // 1..16 whole rows
    float r[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (i < rows) {
        const int q = qs + i;
        const int4 w4 = *reinterpret_cast<const int4*>(tb8 + 16 * (q + (q >> 4)));
        const unsigned w[4] = {static_cast<unsigned>(w4.x), static_cast<unsigned>(w4.y), static_cast<unsigned>(w4.z),
                               static_cast<unsigned>(w4.w)};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const unsigned x = w[k] ^ 0x80008000u;    // exact int16 -> float: bits(2^23 + (s + 32768)) - (2^23 + 32768)
          const float s0 = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7410)) - 8421376.0f;
          const float s1 = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7432)) - 8421376.0f;
          const float p0 = __fmul_rn(s0, s0), p1 = __fmul_rn(s1, s1);
          r[2 * k] = i == 0 ? p0 : __fadd_rn(r[2 * k], p0);
          r[2 * k + 1] = i == 0 ? p1 : __fadd_rn(r[2 * k + 1], p1);
        }
      }
    }
    float v = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    // fold the rounds: round k merges with the stack while the low bits of k are ones (left + right)
    for (int k = round; k & 1; k >>= 1) v = __fadd_rn(sc.v_val[--sp], v);
    __syncwarp();                                   // also: every lane is done with this round's tile
    if (lane == 0) sc.v_val[sp] = v;
    ++sp;
    __syncwarp();
  }
  if (lane == 0)
    *out = static_cast<float>(static_cast<double>(__fmul_rn(sc.v_val[0], 9.313225746154785e-10f)) / static_cast<double>(L));
}

// Persistent CTAs: the leaf table is built once per CTA, every warp then takes clips b0 + warp, b0 + 8 * gridDim.x ...
template <int DT, int kPowWarps>
__global__ void __launch_bounds__(kPowWarps * 32) clip_power_kernel(const void* __restrict__ audio,
                                                                     const long long* __restrict__ offsets,
                                                                     const int* __restrict__ lengths,
                                                                     float* __restrict__ power, const int n_clips,
                                                                     const int aligned) {
  __shared__ PowScratch scratch[kPowWarps];
  __shared__ PowTable table;
  extern __shared__ __align__(16) float pow_tiles[];     // [kPowWarps][2 * kPowPerfTile]
  float (*tiles)[2 * kPowPerfTile] = reinterpret_cast<float (*)[2 * kPowPerfTile]>(pow_tiles);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  static_assert(kPowPerfTile >= 4 * kPowTileGroup && kPowWarps * kPowPerfTile >= 6 * kPowMaxLeaves, "the tiles double as the build buffer");
  if (warp == 0) pow_table_build(table, lengths[min(n_clips - 1, blockIdx.x * kPowWarps)], reinterpret_cast<int*>(&tiles[0][0]), lane);
  __syncthreads();
  for (int b = blockIdx.x * kPowWarps + warp; b < n_clips; b += gridDim.x * kPowWarps) {
    const int L = lengths[b];
    const long long base = offsets[b];
    if (L <= 0) {                                 // np.mean of an empty array is nan
      if (lane == 0) power[b] = __int_as_float(0x7fc00000);
      continue;
    }
    if (table.n_leaves > 0 && table.length == L) {
      if constexpr (DT == ASR_I16) {
        if (table.perfect && aligned && (base & 7) == 0) {
          clip_power_perfect<DT>(audio, base, L, table, scratch[warp], tiles[warp], lane, power + b);
          __syncwarp();
          continue;
        }
      }
      if (aligned && (base & 7) == 0) clip_power_replay_vec<DT>(audio, base, L, table, scratch[warp], tiles[warp], lane, power + b);
      else clip_power_replay<DT>(audio, base, L, table, scratch[warp], lane, power + b);
    } else {
      clip_power_walk<DT>(audio, base, L, scratch[warp], lane, power + b);
    }
    __syncwarp();
  }
}

// Small transfers between MAPPED pinned host memory and device memory as a kernel (loads / stores over PCIe) instead of a
// copy-engine operation: a DMA copy of a few KB queues behind whatever bulk copy the engine of its direction is busy with
// (the 262 MB audio upload of the next batch), which put the 4*B-byte power read-back and the 8*B-byte sigma upload of a step
// behind 4.8 ms of unrelated traffic (scripts/e2e_timeline.py).
__global__ void copy_words_kernel(const unsigned* __restrict__ src, unsigned* __restrict__ dst, const size_t n_words) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n_words; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    dst[i] = src[i];
}

__global__ void snr_sigma_kernel(const float* __restrict__ power, const float snr_db, double* __restrict__ sigma,
                                 const int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float p = power[i];
  const float lg = static_cast<float>(log10(static_cast<double>(p)));      // np.log10(float32)
  const float sdb = __fmul_rn(10.0f, lg);                                    // 10 * ...
  const float ndb = __fsub_rn(sdb, snr_db);                                  // - target_snr_db
  const float t = __fdiv_rn(ndb, 10.0f);                                     // / 10
  const float w = static_cast<float>(pow(10.0, static_cast<double>(t)));    // 10 ** ...
  sigma[i] = static_cast<double>(__fsqrt_rn(w));                             // np.sqrt
}

__device__ __forceinline__ double audio_f64(const void* __restrict__ audio, const int dtype, const long long i) {
  if (dtype == ASR_I16) return static_cast<double>(static_cast<float>(__ldg(reinterpret_cast<const short*>(audio) + i)) * (1.0f / 32768.0f));
  if (dtype == ASR_F32) return static_cast<double>(__ldg(reinterpret_cast<const float*>(audio) + i));
  return __ldg(reinterpret_cast<const double*>(audio) + i);
}

__global__ void __launch_bounds__(256) mix_white_kernel(const void* __restrict__ audio, const int dtype,
                                                        const long long* __restrict__ offsets,
                                                        const int* __restrict__ lengths, const double* __restrict__ z,
                                                        const double* __restrict__ sigma, double* __restrict__ out) {
  const int b = blockIdx.x;
  const int L = lengths[b];
  const long long base = offsets[b];
  const double s = sigma[b];
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < L; i += gridDim.y * blockDim.x)
    out[base + i] = __dadd_rn(audio_f64(audio, dtype, base + i), __dmul_rn(s, __ldg(z + base + i)));
}

// Babble noise (BASELINE configs[1] "white/babble"; no reference implementation - the recipe is SURVEY.md 8(d)):
//   b_i[n] = sum_{k=1..talkers} x_{(i + k*stride) mod B}[n]   (n < L_i; a talker shorter than n contributes nothing)
// in float64 (a sum of at most 6 float32 / int16 values is exact in float64, so the order does not matter), and
//   Pb_i = mean_n b_i[n]^2 in float64 in a FIXED order (256 strided partials per clip, then a binary tree).
// The mix itself is the white-noise formula with z := b and sigma := gain (float64 x + gain*b, two roundings), so the
// fused MFCC launch and asr_mix_white take the stream as it is.  One CTA per clip.
// One CTA per clip walks the clip's chunks of 2048 samples in ascending order (the clips of a launch advance together, so
// the six re-reads of every source chunk hit L2: 280 MB of DRAM reads for 1.5 GB of loads on the C2 batch).  Per chunk the
// 256 threads take 8 consecutive samples each (one vector load per talker; int16 talkers are summed as integers and
// converted once), the chunk's squares are added in the FIXED order the oracle restates - per thread its 8 samples in
// order, then the binary tree (t, t + 128), (t, t + 64), ... (t, t + 1) over the 256 thread sums, then the chunk sums in
// ascending order.  The tree's first three levels pair whole warps: they are taken by warp 0 from shared memory, the
// last five are warp shuffles (floating-point addition is commutative, so only the tree's shape is fixed).
constexpr int kBabbleChunk = 2048;
__global__ void __launch_bounds__(256) babble_stream_kernel(const void* __restrict__ audio, const int dtype,
                                                            const long long* __restrict__ offsets,
                                                            const int* __restrict__ lengths, const int n_clips,
                                                            const int stride, const int talkers,
                                                            double* __restrict__ b_out, double* __restrict__ power_out) {
  __shared__ double s_red[256];
  __shared__ long long s_off[8];
  __shared__ int s_len[8];
  const int i = blockIdx.x;
  const int L = lengths[i];
  const int n_chunks = (L + kBabbleChunk - 1) / kBabbleChunk;
  const long long base = offsets[i];
  const int nt = min(talkers, 8);
  if (threadIdx.x < nt) {
    const int j = static_cast<int>((static_cast<long long>(i) + static_cast<long long>(threadIdx.x + 1) * stride) % n_clips);
    s_off[threadIdx.x] = offsets[j];
    s_len[threadIdx.x] = min(lengths[j], L);
  }
  __syncthreads();
  double total = 0.0;                                  // thread 0: the chunk sums in ascending order
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int n0 = ch * kBabbleChunk + 8 * threadIdx.x;
    double b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int si[8] = {0, 0, 0, 0, 0, 0, 0, 0};          // int16 talkers: the sum is an exact integer, converted once
    for (int k = 0; k < talkers; ++k) {
      long long oj; int Lj;
      if (k < 8) { oj = s_off[k]; Lj = s_len[k]; }
      else {
        const int j = static_cast<int>((static_cast<long long>(i) + static_cast<long long>(k + 1) * stride) % n_clips);
        oj = offsets[j]; Lj = min(lengths[j], L);
      }
      if (n0 + 8 <= Lj && dtype == ASR_I16 && ((oj + n0) & 7) == 0) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(reinterpret_cast<const short*>(audio) + oj + n0));
        const int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          si[2 * e] += static_cast<int>(static_cast<short>(w[e] & 0xFFFF));
          si[2 * e + 1] += w[e] >> 16;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (n0 + e < Lj) b[e] += audio_f64(audio, dtype, oj + n0 + e);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) b[e] += static_cast<double>(si[e]) * (1.0 / 32768.0);   // exact: integers below 2^53 times a power of two
    double acc = 0.0;
    if (n0 + 8 <= L && ((base + n0) & 1) == 0) {     // 16-byte aligned: four vector stores
      double2* o2 = reinterpret_cast<double2*>(b_out + base + n0);
#pragma unroll
      for (int e = 0; e < 4; ++e) o2[e] = make_double2(b[2 * e], b[2 * e + 1]);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc = __dadd_rn(acc, __dmul_rn(b[e], b[e]));
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (n0 + e < L) {
          b_out[base + n0 + e] = b[e];
          acc = __dadd_rn(acc, __dmul_rn(b[e], b[e]));
        }
    }
    s_red[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      const int l = threadIdx.x;
      // levels (t, t+128), (t, t+64), (t, t+32) for column l: e_j = s_red[l + 32 j]
      const double e0 = s_red[l], e1 = s_red[l + 32], e2 = s_red[l + 64], e3 = s_red[l + 96];
      const double e4 = s_red[l + 128], e5 = s_red[l + 160], e6 = s_red[l + 192], e7 = s_red[l + 224];
      double v = __dadd_rn(__dadd_rn(__dadd_rn(e0, e4), __dadd_rn(e2, e6)), __dadd_rn(__dadd_rn(e1, e5), __dadd_rn(e3, e7)));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_down_sync(0xffffffffu, v, o));   // (t, t+16) ... (t, t+1)
      if (l == 0) total = __dadd_rn(total, v);
    }
    __syncthreads();                                   // s_red is rewritten by the next chunk
  }
  if (threadIdx.x == 0) power_out[i] = L > 0 ? total / static_cast<double>(L) : 0.0;
}

__global__ void __launch_bounds__(256) mix_mixture_kernel(const void* __restrict__ audio, const int dtype,
                                                          const long long* __restrict__ offsets,
                                                          const int* __restrict__ lengths, const double* __restrict__ q,
                                                          const double* __restrict__ g, const double p, const double s0,
                                                          const double s1, double* __restrict__ out) {
  const int b = blockIdx.x;
  const int L = lengths[b];
  const long long base = offsets[b];
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < L; i += gridDim.y * blockDim.x) {
    const double sel = (fabs(__ldg(q + base + i)) < p) ? s1 : s0;
    out[base + i] = __dadd_rn(audio_f64(audio, dtype, base + i), __dmul_rn(sel, __ldg(g + base + i)));
  }
}

__global__ void __launch_bounds__(256) mix_rows_kernel(const double* __restrict__ x, const long long n,
                                                       const double* __restrict__ q, const double* __restrict__ g,
                                                       const int mixture, const double p, const double s0,
                                                       const double s1, double* __restrict__ out) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    double sel = s0;
    if (mixture) sel = (fabs(__ldg(q + i)) < p) ? s1 : s0;
    out[i] = __dadd_rn(__ldg(x + i), __dmul_rn(sel, __ldg(g + i)));
  }
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11); counter = element index / 4, key = seed.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&o)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

__global__ void __launch_bounds__(256) randn_kernel(const unsigned long long seed, const unsigned long long first,
                                                    const long long n, double* __restrict__ out) {
  // one thread per group of 4 consecutive GLOBAL indices (aligned to 4), so a value only depends on (seed, index)
  const unsigned long long g0 = first / 4;
  const unsigned long long g1 = (first + n + 3) / 4;
  for (unsigned long long g = g0 + static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; g < g1;
       g += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
    uint32_t r[4];
    philox4x32_10(static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32), 0u, 0u, static_cast<uint32_t>(seed),
                  static_cast<uint32_t>(seed >> 32), r);
    float zf[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float u1 = (static_cast<float>(r[2 * h] >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0,1)
      const float u2 = (static_cast<float>(r[2 * h + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
      const float rad = sqrtf(-2.0f * logf(u1));
      float s, c;
      sincospif(2.0f * u2, &s, &c);
      zf[2 * h] = rad * c;
      zf[2 * h + 1] = rad * s;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned long long idx = g * 4 + j;
      if (idx >= first && idx < first + static_cast<unsigned long long>(n)) out[idx - first] = static_cast<double>(zf[j]);
    }
  }
}

}  // namespace asr

// ------------------------------------------------------------------------------------------------
using namespace asr;

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

template <int NW>
static int launch_clip_power(const void* audio_dev, int32_t dtype, const long long* off, const int32_t* lengths_dev, int32_t n_clips,
                             float* power_dev, int aligned, void* stream) {
  const int blocks = std::min((n_clips + NW - 1) / NW, 148 * (kPowWarpsWide / NW));   // persistent: the SMs' worth of CTAs
  constexpr int smem = NW * 2 * kPowPerfTile * static_cast<int>(sizeof(float));
  static int granted_i16[kMaxDevices] = {0}, granted_f32[kMaxDevices] = {0};
  ASR_CUDA_TRY(ensure_dyn_smem(clip_power_kernel<ASR_I16, NW>, smem, 0, granted_i16));
  ASR_CUDA_TRY(ensure_dyn_smem(clip_power_kernel<ASR_F32, NW>, smem, 0, granted_f32));
  if (dtype == ASR_I16)
    clip_power_kernel<ASR_I16, NW><<<blocks, NW * 32, smem, as_stream(stream)>>>(audio_dev, off, lengths_dev, power_dev, n_clips, aligned);
  else
    clip_power_kernel<ASR_F32, NW><<<blocks, NW * 32, smem, as_stream(stream)>>>(audio_dev, off, lengths_dev, power_dev, n_clips, aligned);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}

extern "C" int asr_clip_power(const void* audio_dev, int32_t dtype, const int64_t* offsets_dev,
                              const int32_t* lengths_dev, int32_t n_clips, float* power_dev, void* stream) {
  if (!audio_dev || !offsets_dev || !lengths_dev || !power_dev || n_clips < 0) {
    set_error("asr_clip_power: null pointer or negative count");
    return ASR_ERR_INVALID;
  }
  if (dtype != ASR_I16 && dtype != ASR_F32) {
    set_error("asr_clip_power: dtype must be ASR_I16 or ASR_F32 (the reference's audio is float32)");
    return ASR_ERR_INVALID;
  }
  if (n_clips == 0) return ASR_OK;
  const long long* off = reinterpret_cast<const long long*>(offsets_dev);
  const int aligned = (reinterpret_cast<uintptr_t>(audio_dev) & 15) == 0 ? 1 : 0;
  // CTA shape: 12 warps x one CTA per SM, or 4 warps x three CTAs per SM - the same warps per SM; the narrow shape needs a
  // third of the shared memory per CTA, so its CTAs find room beside the tail kernels of a step when the pass runs on a side
  // stream (asr_b200.pipeline).  ASR_B200_POW_WARPS = 12 | 4 overrides the default.
  static const int shape = [] { const char* e = std::getenv("ASR_B200_POW_WARPS"); return e ? std::atoi(e) : kPowDefaultWarps; }();
  if (shape == kPowWarpsNarrow) return launch_clip_power<kPowWarpsNarrow>(audio_dev, dtype, off, lengths_dev, n_clips, power_dev, aligned, stream);
  return launch_clip_power<kPowWarpsWide>(audio_dev, dtype, off, lengths_dev, n_clips, power_dev, aligned, stream);
}

extern "C" int asr_copy_mapped(const void* src, void* dst, size_t bytes, void* stream) {
  if ((!src || !dst) && bytes != 0) { set_error("asr_copy_mapped: null pointer"); return ASR_ERR_INVALID; }
  if ((bytes & 3) != 0 || ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 3) != 0) {
    set_error("asr_copy_mapped: pointers and byte count must be multiples of 4");
    return ASR_ERR_INVALID;
  }
  if (bytes == 0) return ASR_OK;
  const size_t n = bytes / 4;
  const int blocks = static_cast<int>(std::min<size_t>((n + 255) / 256, 64));
  copy_words_kernel<<<blocks, 256, 0, as_stream(stream)>>>(static_cast<const unsigned*>(src), static_cast<unsigned*>(dst), n);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}

extern "C" int asr_snr_sigma(const float* power_dev, float target_snr_db, double* sigma_dev, int32_t n_clips,
                             void* stream) {
  if (!power_dev || !sigma_dev || n_clips < 0) {
    set_error("asr_snr_sigma: null pointer or negative count");
    return ASR_ERR_INVALID;
  }
  if (n_clips == 0) return ASR_OK;
  snr_sigma_kernel<<<(n_clips + 255) / 256, 256, 0, as_stream(stream)>>>(power_dev, target_snr_db, sigma_dev, n_clips);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}

// Host side of the SNR chain (VDR/attacks.py:235-241 executed under numpy >= 2 scalar rules, every step float32):
//   db = 10 * log10(P) ; ndb = db - snr ; w = 10 ** (ndb / 10) ; sigma = sqrt(w)
// numpy evaluates the scalar `10 ** x` with libm's powf and the other steps with correctly rounded float32 arithmetic,
// so those are restated here against the same libm.  np.log10 is NOT libm on every host (numpy dispatches to its own
// SIMD kernels on AVX-512 machines; about half of all inputs then differ from glibc's log10f in the last bit), which is
// why the caller may hand in log10(P) as numpy computed it (`log10_power_host`); NULL = glibc's log10f.
// Pure host code: no CUDA call, no device memory.  Built with -ffp-contract=off (no fused multiply-subtract).
extern "C" int asr_snr_sigma_host(const float* power_host, const float* log10_power_host, float target_snr_db,
                                  double* sigma_host, int32_t n_clips) {
  if (!power_host || !sigma_host || n_clips < 0) {
    set_error("asr_snr_sigma_host: null pointer or negative count");
    return ASR_ERR_INVALID;
  }
  for (int32_t i = 0; i < n_clips; ++i) {
    const volatile float lg = log10_power_host ? log10_power_host[i] : log10f(power_host[i]);
    const volatile float db = 10.0f * lg;
    const volatile float ndb = db - target_snr_db;
    const volatile float q = ndb / 10.0f;
    const volatile float w = powf(10.0f, q);
    sigma_host[i] = static_cast<double>(sqrtf(w));
  }
  return ASR_OK;
}

extern "C" int asr_mix_white(const void* audio_dev, int32_t dtype, const int64_t* offsets_dev,
                             const int32_t* lengths_dev, int32_t n_clips, const double* z_dev,
                             const double* sigma_dev, double* out_dev, void* stream) {
  if (!audio_dev || !offsets_dev || !lengths_dev || !z_dev || !sigma_dev || !out_dev || n_clips < 0 ||
      dtype < ASR_I16 || dtype > ASR_F64) {
    set_error("asr_mix_white: invalid argument");
    return ASR_ERR_INVALID;
  }
  if (n_clips == 0) return ASR_OK;
  mix_white_kernel<<<dim3(n_clips, 8), 256, 0, as_stream(stream)>>>(
      audio_dev, dtype, reinterpret_cast<const long long*>(offsets_dev), lengths_dev, z_dev, sigma_dev, out_dev);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}

extern "C" size_t asr_babble_workspace_bytes(int32_t n_clips, int32_t max_length) {
  if (n_clips <= 0 || max_length < 0) return 0;
  const size_t chunks = static_cast<size_t>(std::max(1, (max_length + kBabbleChunk - 1) / kBabbleChunk));
  return ((sizeof(int) * static_cast<size_t>(n_clips) + 255) & ~static_cast<size_t>(255)) + sizeof(double) * chunks * n_clips;
}

extern "C" int asr_babble_stream(const void* audio_dev, int32_t dtype, const int64_t* offsets_dev, const int32_t* lengths_dev,
                                 int32_t n_clips, int32_t max_length, int32_t stride, int32_t talkers, double* babble_dev,
                                 double* power_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
  if (!audio_dev || !offsets_dev || !lengths_dev || !babble_dev || !power_dev || n_clips < 0 || talkers < 1 || stride < 1 ||
      max_length < 0 || dtype < ASR_I16 || dtype > ASR_F64) {
    set_error("asr_babble_stream: invalid argument");
    return ASR_ERR_INVALID;
  }
  if (n_clips == 0) return ASR_OK;
  if (!workspace_dev || workspace_bytes < asr_babble_workspace_bytes(n_clips, max_length) ||
      (reinterpret_cast<uintptr_t>(workspace_dev) & 7)) {
    set_error("asr_babble_stream: workspace too small (asr_babble_workspace_bytes) or misaligned; it must be zeroed once");
    return ASR_ERR_INVALID;
  }
  // (the workspace of earlier versions - chunk partials and arrival counters - is no longer used: a clip's chunks are summed
  //  inside its CTA; the argument stays in the ABI and is still validated)
  babble_stream_kernel<<<n_clips, 256, 0, as_stream(stream)>>>(
      audio_dev, dtype, reinterpret_cast<const long long*>(offsets_dev), lengths_dev, n_clips, stride, talkers, babble_dev, power_dev);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}

extern "C" int asr_mix_mixture(const void* audio_dev, int32_t dtype, const int64_t* offsets_dev,
                               const int32_t* lengths_dev, int32_t n_clips, const double* q_dev, const double* g_dev,
                               double p, double sigma0, double sigma1, double* out_dev, void* stream) {
  if (!audio_dev || !offsets_dev || !lengths_dev || !q_dev || !g_dev || !out_dev || n_clips < 0 ||
      dtype < ASR_I16 || dtype > ASR_F64) {
    set_error("asr_mix_mixture: invalid argument");
    return ASR_ERR_INVALID;
  }
  if (n_clips == 0) return ASR_OK;
  mix_mixture_kernel<<<dim3(n_clips, 8), 256, 0, as_stream(stream)>>>(
      audio_dev, dtype, reinterpret_cast<const long long*>(offsets_dev), lengths_dev, q_dev, g_dev, p, sigma0, sigma1,
      out_dev);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}

extern "C" int asr_mix_rows_white(const double* x_dev, int64_t n, const double* z_dev, double sigma, double* out_dev,
                                  void* stream) {
  if (!x_dev || !z_dev || !out_dev || n < 0) {
    set_error("asr_mix_rows_white: invalid argument");
    return ASR_ERR_INVALID;
  }
  if (n == 0) return ASR_OK;
  const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, 148 * 16));
  mix_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x_dev, n, nullptr, z_dev, 0, 0.0, sigma, sigma, out_dev);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}

extern "C" int asr_mix_rows_mixture(const double* x_dev, int64_t n, const double* q_dev, const double* g_dev, double p,
                                    double sigma0, double sigma1, double* out_dev, void* stream) {
  if (!x_dev || !q_dev || !g_dev || !out_dev || n < 0) {
    set_error("asr_mix_rows_mixture: invalid argument");
    return ASR_ERR_INVALID;
  }
  if (n == 0) return ASR_OK;
  const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, 148 * 16));
  mix_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x_dev, n, q_dev, g_dev, 1, p, sigma0, sigma1, out_dev);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}

extern "C" int asr_randn_f64(uint64_t seed, uint64_t first_index, int64_t n, double* out_dev, void* stream) {
  if (!out_dev || n < 0) {
    set_error("asr_randn_f64: invalid argument");
    return ASR_ERR_INVALID;
  }
  if (n == 0) return ASR_OK;
  const int64_t groups = (n + 3) / 4 + 1;
  const int blocks = static_cast<int>(std::min<int64_t>((groups + 255) / 256, 148 * 16));
  randn_kernel<<<blocks, 256, 0, as_stream(stream)>>>(seed, first_index, n, out_dev);
  ASR_CUDA_TRY(cudaGetLastError());
  return ASR_OK;
}
