// C-ABI of libasr_b200: plan construction (host tables in float64 -> float32 device blob),
// the fused-MFCC launch and the host-buffer pipeline.  See include/asr_b200.h for the contract.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <dlfcn.h>
#include <numeric>
#include <cuda_fp16.h>
#include "common.cuh"

namespace asr {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int cuda_fail(cudaError_t e, const char* what) {
  g_last_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
  return ASR_ERR_CUDA;
}

static inline int round4(int x) { return (x + 3) & ~3; }
// smallest multiple of 4 >= x whose quarter is odd: float4 rows at this pitch are bank-conflict free
static inline int pitch_odd4(int x) {
  int p = round4(x);
  if (((p / 4) & 1) == 0) p += 4;
  return p;
}

static const double kPi = 3.141592653589793238462643383279502884;

// ---- librosa.core.convert (Slaney mel scale) ------------------------------------------------------
static double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
  const double logstep = std::log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
  const double logstep = std::log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}
static std::vector<double> linspace(double a, double b, int n) {
  std::vector<double> v(n);
  if (n == 1) { v[0] = a; return v; }
  const double step = (b - a) / (n - 1);
  for (int i = 0; i < n; ++i) v[i] = a + i * step;
  v[n - 1] = b;
  return v;
}

// librosa.filters.mel(htk=False, norm='slaney', dtype=float32), row-major (n_mels, n_bins)
static std::vector<float> mel_dense(const asr_mfcc_params& p, int n_bins, std::vector<double>* mel_f_out = nullptr,
                                    std::vector<double>* fftfreqs_out = nullptr) {
  const double fmax = p.fmax > 0 ? static_cast<double>(p.fmax) : p.sr / 2.0;
  std::vector<double> fftfreqs(n_bins);
  if (p.fftfreq_mode == ASR_FFTFREQ_LINSPACE) {
    fftfreqs = linspace(0.0, p.sr / 2.0, n_bins);
  } else {
    const double val = 1.0 / (p.n_fft * (1.0 / p.sr));
    for (int k = 0; k < n_bins; ++k) fftfreqs[k] = k * val;
  }
  std::vector<double> mel_f = linspace(hz_to_mel(p.fmin), hz_to_mel(fmax), p.n_mels + 2);
  for (double& m : mel_f) m = mel_to_hz(m);
  std::vector<float> w(static_cast<size_t>(p.n_mels) * n_bins, 0.0f);
  for (int i = 0; i < p.n_mels; ++i) {
    const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
    const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
    for (int k = 0; k < n_bins; ++k) {
      const double lower = -(mel_f[i] - fftfreqs[k]) / fd0;
      const double upper = (mel_f[i + 2] - fftfreqs[k]) / fd1;
      const float tri = static_cast<float>(std::max(0.0, std::min(lower, upper)));   // stored float32 ...
      w[static_cast<size_t>(i) * n_bins + k] = static_cast<float>(static_cast<double>(tri) * enorm);  // ... *= enorm
    }
  }
  if (mel_f_out) *mel_f_out = mel_f;
  if (fftfreqs_out) *fftfreqs_out = fftfreqs;
  return w;
}

// scipy.signal.savgol_coeffs(width, polyorder=order, deriv=order) as correlation taps
static bool savgol_taps(int width, int order, double* taps) {
  const int h = width / 2, m = order + 1;
  double A[3][3] = {{0}}, inv[3][3] = {{0}};
  for (int x = -h; x <= h; ++x)
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) A[i][j] += std::pow(static_cast<double>(x), i + j);
  // Gauss-Jordan inverse of the (m x m) normal matrix
  double aug[3][6];
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < 2 * m; ++j) aug[i][j] = j < m ? A[i][j] : (j - m == i ? 1.0 : 0.0);
  for (int c = 0; c < m; ++c) {
    int piv = c;
    for (int r = c + 1; r < m; ++r)
      if (std::fabs(aug[r][c]) > std::fabs(aug[piv][c])) piv = r;
    if (std::fabs(aug[piv][c]) < 1e-300) return false;
    for (int j = 0; j < 2 * m; ++j) std::swap(aug[c][j], aug[piv][j]);
    const double d = aug[c][c];
    for (int j = 0; j < 2 * m; ++j) aug[c][j] /= d;
    for (int r = 0; r < m; ++r)
      if (r != c) {
        const double f = aug[r][c];
        for (int j = 0; j < 2 * m; ++j) aug[r][j] -= f * aug[c][j];
      }
  }
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) inv[i][j] = aug[i][m + j];
  double fact = 1.0;
  for (int i = 2; i <= order; ++i) fact *= i;
  for (int x = -h; x <= h; ++x) {
    double t = 0.0;
    for (int j = 0; j < m; ++j) t += inv[order][j] * std::pow(static_cast<double>(x), j);
    taps[x + h] = fact * t;
  }
  return true;
}


// ---- cepstra tables (cepstra_*_kernel, shared by every path that goes through the log-mel workspace) ----
// DCT (lifter folded in) transposed to [lm_pitch][4*NC4] (zero rows / columns as padding), then the delta taps.
static bool build_cepstra_tables(asr_plan* pl) {
  const asr_mfcc_params& p = pl->prm;
  const int n_mels = p.n_mels;
  pl->fr_lm_pitch = round4(n_mels);
  if (p.n_mfcc > 40) return true;
  const int nc4 = (p.n_mfcc + 3) / 4;
  std::vector<float> blob;
  auto put_f = [&](const float* src, size_t n) {
    const int off = static_cast<int>(blob.size());
    blob.insert(blob.end(), src, src + n);
    blob.resize(round4(static_cast<int>(blob.size())), 0.0f);
    return off;
  };
  std::vector<float> dct_t(static_cast<size_t>(pl->fr_lm_pitch) * 4 * nc4, 0.0f);
  for (int c = 0; c < p.n_mfcc; ++c)
    for (int j = 0; j < n_mels; ++j) dct_t[static_cast<size_t>(j) * 4 * nc4 + c] = pl->h_dct[static_cast<size_t>(c) * n_mels + j];
  pl->h_dct_t = dct_t;
  const int dct_off = put_f(dct_t.data(), dct_t.size());
  const int taps_off = put_f(pl->h_taps.data(), pl->h_taps.size());
  pl->cep_tab_f4 = static_cast<int>(blob.size() / 4);
  pl->cep_off_taps = taps_off - dct_off;
  pl->cep_off_cbuf = 4 * pl->cep_tab_f4;
  pl->cep_smem_bytes = 4 * (pl->cep_off_cbuf + (p.delta_orders > 0 ? p.n_mfcc * 129 : 0));
  if (cudaMalloc(reinterpret_cast<void**>(&pl->cep_dev), blob.size() * sizeof(float)) != cudaSuccess) return false;
  if (cudaMemcpy(pl->cep_dev, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return false;
  return true;
}

// ---- tables of the block-pipelined path (frames_kernel.cu), n_fft = 512 ------------------------------
// Mel bank in "segment" form: the Slaney triangles overlap only their neighbours, so a bin between the
// peaks of filters seg-1 and seg feeds the falling slope of seg-1 and the rising slope of seg.  Bins are
// cut into kFrWarps ranges (one per warp); a range x segment intersection is a "piece" with two partial
// sums (fall, rise); a filter is the sum of a short list of partials ("refs"), in ascending-bin order.
static bool build_frames_tables(asr_plan* pl, const std::vector<double>& mel_f, const std::vector<double>& fftfreqs,
                                const std::vector<float>& twp, const std::vector<float>& twu) {
  const asr_mfcc_params& p = pl->prm;
  const int n_bins = pl->n_bins, n_mels = p.n_mels;
  pl->fr_ok = 0;
  if (p.n_fft != 512 || !pl->fft_path || (p.hop_length & 1) || p.n_mfcc > 40) return true;
  std::vector<int> seg(n_bins);
  for (int k = 0; k < n_bins; ++k) {
    int s = 0;
    while (s < n_mels && mel_f[s + 1] <= fftfreqs[k]) ++s;
    seg[k] = s;
  }
  for (int i = 0; i < n_mels; ++i)
    for (int k = 0; k < n_bins; ++k)
      if (pl->h_mel_dense[static_cast<size_t>(i) * n_bins + k] != 0.0f && i != seg[k] - 1 && i != seg[k]) return true;   // not neighbour-only
  auto W = [&](int i, int k) { return (i >= 0 && i < n_mels) ? pl->h_mel_dense[static_cast<size_t>(i) * n_bins + k] : 0.0f; };
  // pieces = whole segments (bins [k0, k1) of segment s), read as float4 groups starting at k0 & ~3
  struct Piece { int k0, k1, seg, nq; };
  std::vector<Piece> pieces;
  int total_q = 0;
  for (int sg = 0; sg <= n_mels; ++sg) {
    int k0 = -1, k1 = -1;
    for (int k = 0; k < n_bins; ++k)
      if (seg[k] == sg) { if (k0 < 0) k0 = k; k1 = k + 1; }
    Piece pc{0, 0, sg, 0};
    if (k0 >= 0) { pc.k0 = k0; pc.k1 = k1; pc.nq = (k1 - (k0 & ~3) + 3) / 4; }
    pieces.push_back(pc);                      // empty segments keep a piece: their partials must be written (zeros)
    total_q += 2 * ((pc.nq + 1) / 2);
  }
  // contiguous groups of segments per warp, balanced by float4 groups (+2 per segment of fixed cost)
  std::vector<int> wrange(2 * kFrWarps, 0);
  {
    const int n_work = kFrMelWarps;             // the last warp assembles block descriptors instead
    const double target = (total_q + 2.0 * pieces.size()) / n_work;
    size_t pi = 0;
    double acc = 0.0;
    for (int w = 0; w < n_work; ++w) {
      wrange[2 * w] = static_cast<int>(pi);
      const double goal = target * (w + 1);
      while (pi < pieces.size() && (w == n_work - 1 || acc + 0.5 * (2 * ((pieces[pi].nq + 1) / 2) + 2) <= goal)) {
        acc += 2 * ((pieces[pi].nq + 1) / 2) + 2;
        ++pi;
      }
      wrange[2 * w + 1] = static_cast<int>(pi) - wrange[2 * w];
    }
  }
  std::vector<float> wtab;                    // per float4 group of bins: (fall, rise) x 4 = two float4
  std::vector<int> ptab;                      // int4 per piece: first bin (multiple of 4), PAIRS of groups, weight offset (float4), fall-partial offset
  for (const Piece& pc : pieces) {
    const int ka = pc.k0 & ~3;
    const int nq2 = (pc.nq + 1) / 2;          // the kernel takes the groups two at a time; a zero-weight group pads odd counts
    ptab.push_back(ka); ptab.push_back(nq2); ptab.push_back(static_cast<int>(wtab.size() / 4)); ptab.push_back(2 * pc.seg * 33);
    for (int i = 0; i < 8 * nq2; ++i) {
      const int k = ka + i;
      const bool in = k >= pc.k0 && k < pc.k1;
      wtab.push_back(in ? W(pc.seg - 1, k) : 0.0f);
      wtab.push_back(in ? W(pc.seg, k) : 0.0f);
    }
  }
  std::vector<int> frange(2, 0), refs(1, 0);  // (unused by the segment form: filter j = part[2j+1] + part[2j+2])
  pl->fr_n_refs = 2 * (n_mels + 1);
  pl->fr_s_pitch = round4(n_bins + 7);        // float4 reads of pairs of 4-bin groups; pitch = 4 (mod 32) floats: lanes <-> frames conflict-free
  if ((pl->fr_s_pitch % 32) != 4) pl->fr_s_pitch += (4 - pl->fr_s_pitch % 32 + 32) % 32;
  pl->fr_xb_stride = pl->frame_stride;
  pl->fr_lm_pitch = round4(n_mels);
  // ---- blob: [frames-kernel tables][cepstra tables] ----
  std::vector<float> blob;
  auto put_f = [&](const float* src, size_t n) {
    const int off = static_cast<int>(blob.size());
    blob.insert(blob.end(), src, src + n);
    blob.resize(round4(static_cast<int>(blob.size())), 0.0f);
    return off;
  };
  auto put_i = [&](const int* src, size_t n) {
    const int off = static_cast<int>(blob.size());
    blob.resize(blob.size() + n);
    if (n) std::memcpy(blob.data() + off, src, n * sizeof(int));
    blob.resize(round4(static_cast<int>(blob.size())), 0.0f);
    return off;
  };
  pl->fr_off_window = put_f(pl->h_window.data(), pl->h_window.size());
  {
    std::vector<float> wi(pl->h_window);
    for (float& v : wi) v *= (1.0f / 32768.0f);
    pl->fr_off_window_i16 = put_f(wi.data(), wi.size());
  }
  pl->fr_off_twp = put_f(twp.data(), twp.size());
  pl->fr_off_twu = put_f(twu.data(), twu.size());
  pl->fr_off_wtab = put_f(wtab.data(), wtab.size());
  pl->fr_off_pieces = put_i(ptab.data(), ptab.size());
  pl->fr_off_wrange = put_i(wrange.data(), wrange.size());
  pl->fr_off_frange = put_i(frange.data(), frange.size());
  pl->fr_off_refs = put_i(refs.data(), refs.size());
  pl->fr_blob_f4 = static_cast<int>(blob.size() / 4);
  // cepstra tables: DCT (lifter folded in) transposed to [lm_pitch][4*NC4] (zero rows / columns as padding), delta taps
  const int nc4 = (p.n_mfcc + 3) / 4;
  std::vector<float> dct_t(static_cast<size_t>(pl->fr_lm_pitch) * 4 * nc4, 0.0f);
  for (int c = 0; c < p.n_mfcc; ++c)
    for (int j = 0; j < n_mels; ++j) dct_t[static_cast<size_t>(j) * 4 * nc4 + c] = pl->h_dct[static_cast<size_t>(c) * n_mels + j];
  pl->cep_blob_f4 = pl->fr_blob_f4;
  pl->h_dct_t = dct_t;
  const int dct_off = put_f(dct_t.data(), dct_t.size());
  const int taps_off = put_f(pl->h_taps.data(), pl->h_taps.size());
  pl->cep_tab_f4 = static_cast<int>(blob.size() / 4) - pl->cep_blob_f4;
  pl->cep_off_taps = taps_off - dct_off;
  pl->cep_off_cbuf = 4 * pl->cep_tab_f4;
  pl->cep_smem_bytes = 4 * (pl->cep_off_cbuf + (p.delta_orders > 0 ? p.n_mfcc * 129 : 0));
  if (cudaMalloc(reinterpret_cast<void**>(&pl->fr_blob_dev), blob.size() * sizeof(float)) != cudaSuccess) return false;
  if (cudaMemcpy(pl->fr_blob_dev, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return false;
  pl->fr_ok = 1;
  return true;
}

// ---- tables of the TMA-staged path (tile_kernel.cu), n_fft = 512 ---------------------------------------
// The same segment form of the mel bank as above, as a flat "program": bins are taken four at a time (one float4 of the
// spectrum row); a step belongs to one segment and carries (fall, rise) weights for its 4 bins (zero outside the segment,
// so a float4 that straddles a boundary appears in two steps).  The steps are dealt out in order to `n_vw` "virtual warps"
// (groups of FR lanes, FR = frames per block); where a share ends inside a segment the segment is cut into pieces, each
// with its own pair of partial rows (row (seg * npc + piece) * 2 = fall, + 1 = rise); filter j = sum of the rise rows of
// segment j and the fall rows of j+1.  With two virtual warps per hardware warp (FR = 16) the shares of a pair are padded
// with empty steps to the same length.
static bool build_tile_tables(asr_plan* pl, int cfg, int n_vw, bool pair_equal, const std::vector<double>& mel_f,
                              const std::vector<double>& fftfreqs, const std::vector<float>& twp, const std::vector<float>& twu) {
  const asr_mfcc_params& p = pl->prm;
  const int fr = cfg == 1 ? 16 : 32;                       // frames per block of this kernel shape
  const int n_bins = pl->n_bins, n_mels = p.n_mels;
  asr_plan::TileTables& tt = pl->tl[cfg];
  std::vector<int> seg(n_bins);
  for (int k = 0; k < n_bins; ++k) {
    int sg = 0;
    while (sg < n_mels && mel_f[sg + 1] <= fftfreqs[k]) ++sg;
    seg[k] = sg;
  }
  auto W = [&](int i, int k) { return (i >= 0 && i < n_mels) ? pl->h_mel_dense[static_cast<size_t>(i) * n_bins + k] : 0.0f; };
  struct Step { int q, seg, piece, last; };
  std::vector<Step> steps;
  for (int sg = 0; sg <= n_mels; ++sg) {
    int k0 = -1, k1 = -1;
    for (int k = 0; k < n_bins; ++k)
      if (seg[k] == sg && (W(sg - 1, k) != 0.0f || W(sg, k) != 0.0f)) { if (k0 < 0) k0 = k; k1 = k + 1; }
    if (k0 < 0) continue;                                   // no bin of this segment carries weight: its partial rows stay 0
    for (int q = k0 / 4; q <= (k1 - 1) / 4; ++q) steps.push_back({q, sg, 0, 0});
  }
  const int ns = static_cast<int>(steps.size());
  if (ns == 0) return true;
  std::vector<std::vector<Step>> share(16);
  std::vector<int> n_pieces(n_mels + 1, 0);
  for (int w = 0; w < n_vw; ++w) {
    const int a = static_cast<int>(static_cast<long long>(w) * ns / n_vw);
    const int b = static_cast<int>(static_cast<long long>(w + 1) * ns / n_vw);
    for (int i = a; i < b; ++i) {
      Step st = steps[i];
      if (i == a || steps[i].seg != steps[i - 1].seg) ++n_pieces[st.seg];
      st.piece = n_pieces[st.seg] - 1;
      st.last = (i + 1 == b || steps[i + 1].seg != st.seg) ? 1 : 0;
      share[w].push_back(st);
    }
  }
  (void)pair_equal;
  int npc = 1;
  for (int sg = 0; sg <= n_mels; ++sg) npc = std::max(npc, n_pieces[sg]);
  if (npc > kTlMaxPieces) return true;
  // per virtual warp a list of pieces (int4: byte offset of the first float4 in the S row, steps, byte offset of the
  // fall partial row, index of the first weight float4); the steps of a piece read consecutive float4 of the row
  std::vector<float> wtab;
  std::vector<int> stab, wrange(2 * 16, 0);
  for (int w = 0; w < 16; ++w) {
    wrange[2 * w] = static_cast<int>(stab.size() / 4);
    int n_pc = 0;
    for (size_t i = 0; i < share[w].size(); ++i) {
      const Step& st = share[w][i];
      if (i == 0 || share[w][i - 1].last) {               // a piece starts here
        stab.push_back(16 * st.q);
        stab.push_back(0);
        stab.push_back((st.seg * npc + st.piece) * 2 * fr * 4);
        stab.push_back(static_cast<int>(wtab.size() / 4));
        ++n_pc;
      }
      stab[stab.size() - 3] += 1;
      for (int e = 0; e < 4; ++e) {
        const int k = 4 * st.q + e;
        const bool in = k < n_bins && seg[k] == st.seg;
        wtab.push_back(in ? 0.25f * W(st.seg - 1, k) : 0.0f);       // the spectrum row of this path holds 4|X|^2 (exact scaling)
        wtab.push_back(in ? 0.25f * W(st.seg, k) : 0.0f);
      }
    }
    wrange[2 * w + 1] = n_pc;
  }
  tt.npc = npc;
  tt.npart = (n_mels + 1) * npc * 2;
  tt.nsteps = static_cast<int>(stab.size() / 4);
  std::vector<float> blob;
  auto put_f = [&](const float* src, size_t n) {
    const int off = static_cast<int>(blob.size());
    blob.insert(blob.end(), src, src + n);
    blob.resize(round4(static_cast<int>(blob.size())), 0.0f);
    return off;
  };
  auto put_i = [&](const int* src, size_t n) {
    const int off = static_cast<int>(blob.size());
    blob.resize(blob.size() + n);
    if (n) std::memcpy(blob.data() + off, src, n * sizeof(int));
    blob.resize(round4(static_cast<int>(blob.size())), 0.0f);
    return off;
  };
  {
    // second-pass lane twiddles of tile_fft512 (dft16_twisted): exp(-2 pi i (l + 16 j) / (16 m)) for (m, j) =
    // (2,0) (4,0) (8,0) (8,1) (16,0..3), lane l = 0..15: 8 complex values per lane
    (void)twp;
    static const int mm[8] = {2, 4, 8, 8, 16, 16, 16, 16}, jj[8] = {0, 0, 0, 1, 0, 1, 2, 3};
    std::vector<float> twl(16 * 16);
    for (int l = 0; l < 16; ++l)
      for (int e = 0; e < 8; ++e) {
        const double a = -2.0 * M_PI * (l + 16 * jj[e]) / (16.0 * mm[e]);
        twl[16 * l + 2 * e] = static_cast<float>(std::cos(a));
        twl[16 * l + 2 * e + 1] = static_cast<float>(std::sin(a));
      }
    tt.off_twp = put_f(twl.data(), twl.size());
  }
  {
    std::vector<float> twu2(twu);
    for (float& v : twu2) v *= 2.0f;                       // the tile kernel unpacks 2X (exact scaling)
    tt.off_twu = put_f(twu2.data(), twu2.size());
  }
  tt.off_wtab = put_f(wtab.data(), wtab.size());
  tt.off_steps = put_i(stab.data(), stab.size());
  tt.off_wrange = put_i(wrange.data(), wrange.size());
  tt.blob_f4 = static_cast<int>(blob.size() / 4);           // copied to shared memory; one of the two windows follows it there
  tt.off_window = put_f(pl->h_window.data(), pl->h_window.size());
  {
    std::vector<float> wi(pl->h_window);
    for (float& v : wi) v *= (1.0f / 32768.0f);
    tt.off_window_i16 = put_f(wi.data(), wi.size());
  }
  if (cudaMalloc(reinterpret_cast<void**>(&tt.blob_dev), blob.size() * sizeof(float)) != cudaSuccess) return false;
  if (cudaMemcpy(tt.blob_dev, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return false;
  return true;
}

static bool build_tile_tables_all(asr_plan* pl, const std::vector<double>& mel_f, const std::vector<double>& fftfreqs,
                                  const std::vector<float>& twp, const std::vector<float>& twu) {
  const asr_mfcc_params& p = pl->prm;
  pl->tl_ok = 0;
  if (!pl->fr_ok || (p.hop_length % 8) != 0 || (pl->pad % 8) != 0 || p.preemph != 0.0f) return true;
  if (!build_tile_tables(pl, 0, 16, false, mel_f, fftfreqs, twp, twu)) return false;     // 16 main warps, 32 frames per block
  if (!build_tile_tables(pl, 1, 16, true, mel_f, fftfreqs, twp, twu)) return false;      // 8 main warps, 16 frames per block
  pl->tl_ok = (pl->tl[0].blob_dev && pl->tl[1].blob_dev) ? 1 : 0;
  return true;
}

// Mel bank for the lanes <-> frames mel stage of the tensor-core kernels (tc_kernel.cu, tcdft_kernel.cu): segment form
// (a bin between the peaks of filters s-1 and s feeds the falling slope of s-1 and the rising slope of s), steps of 4 bins,
// dealt out to the 4 warps of a TMEM lane quarter; a filter whose terms all lie in one share is finished in registers
// ("direct"), the others go through boundary slots.  `scale` undoes the scaling of the power spectrum in TMEM.
struct LaneMel { float* blob_dev; int blob_f4, off_wtab, off_pieces, off_wrange, off_bnd, n_bnd, n_slots, ok; };
static bool build_lane_mel(asr_plan* pl, const std::vector<double>& mel_f, const std::vector<double>& fftfreqs, const float scale,
                           LaneMel* lm) {
  const asr_mfcc_params& p = pl->prm;
  const int n_bins = pl->n_bins, n_mels = p.n_mels;
  lm->ok = 0; lm->blob_dev = nullptr;
  if (n_mels > 254) return true;
  for (int i = 0; i < n_mels; ++i) {                       // neighbour-only check (Slaney triangles overlap their neighbours only)
    for (int k = 0; k < n_bins; ++k) {
      int sg = 0;
      while (sg < n_mels && mel_f[sg + 1] <= fftfreqs[k]) ++sg;
      if (pl->h_mel_dense[static_cast<size_t>(i) * n_bins + k] != 0.0f && i != sg - 1 && i != sg) return true;
    }
  }
  // ---- mel bank: steps (4 bins of one segment), 4 contiguous shares ----
  std::vector<int> seg(n_bins);
  for (int k = 0; k < n_bins; ++k) {
    int sg = 0;
    while (sg < n_mels && mel_f[sg + 1] <= fftfreqs[k]) ++sg;
    seg[k] = sg;
  }
  auto W = [&](int i, int k) { return (i >= 0 && i < n_mels) ? pl->h_mel_dense[static_cast<size_t>(i) * n_bins + k] : 0.0f; };
  struct Step { int q, seg; };
  std::vector<Step> steps;
  for (int sg = 0; sg <= n_mels; ++sg) {
    int k0 = -1, k1 = -1;
    for (int k = 0; k < n_bins; ++k)
      if (seg[k] == sg && (W(sg - 1, k) != 0.0f || W(sg, k) != 0.0f)) { if (k0 < 0) k0 = k; k1 = k + 1; }
    if (k0 < 0) continue;
    for (int q = k0 / 4; q <= (k1 - 1) / 4; ++q) steps.push_back({q, sg});
  }
  const int ns = static_cast<int>(steps.size());
  if (ns < 4) return true;
  struct Piece { int q0, nsteps, w0, filter, group; };
  std::vector<Piece> pieces;
  std::vector<float> wtab;
  std::vector<int> wrange(2 * 4, 0);
  std::vector<std::vector<int>> contrib(n_mels);            // per filter: indices of the pieces that emit it
  for (int g = 0; g < 4; ++g) {
    const int a = static_cast<int>(static_cast<long long>(g) * ns / 4), b = static_cast<int>(static_cast<long long>(g + 1) * ns / 4);
    wrange[2 * g] = static_cast<int>(pieces.size());
    int i = a;
    for (int sg = steps[a].seg; sg <= steps[b - 1].seg; ++sg) {
      Piece pc{0, 0, static_cast<int>(wtab.size() / 4), sg - 1, g};
      if (i < b && steps[i].seg == sg) {
        pc.q0 = steps[i].q;
        while (i < b && steps[i].seg == sg) {
          if (steps[i].q != pc.q0 + pc.nsteps) return true;            // steps of a segment are consecutive float4 groups
          for (int e = 0; e < 4; ++e) {
            const int k = 4 * steps[i].q + e;
            const bool in = k < n_bins && seg[k] == sg;
            wtab.push_back(in ? scale * W(sg - 1, k) : 0.0f);
            wtab.push_back(in ? scale * W(sg, k) : 0.0f);
          }
          ++pc.nsteps; ++i;
        }
      }
      pieces.push_back(pc);
    }
    pieces.push_back(Piece{0, 0, static_cast<int>(wtab.size() / 4), steps[b - 1].seg, g});   // the pending rise sum
    wrange[2 * g + 1] = static_cast<int>(pieces.size()) - wrange[2 * g];
  }
  for (size_t i = 0; i < pieces.size(); ++i)
    if (pieces[i].filter >= 0 && pieces[i].filter < n_mels) contrib[pieces[i].filter].push_back(static_cast<int>(i));
  std::vector<int> ptab(4 * pieces.size(), 0), bnd;
  int n_slots = 0;
  for (size_t i = 0; i < pieces.size(); ++i) {
    ptab[4 * i] = pieces[i].q0; ptab[4 * i + 1] = pieces[i].nsteps; ptab[4 * i + 2] = pieces[i].w0; ptab[4 * i + 3] = 0;
  }
  for (int f = 0; f < n_mels; ++f) {
    const std::vector<int>& cl = contrib[f];
    if (cl.size() == 1) { ptab[4 * cl[0] + 3] = f + 1; continue; }
    bnd.push_back(f); bnd.push_back(n_slots); bnd.push_back(static_cast<int>(cl.size())); bnd.push_back(0);
    for (int pi : cl) { ptab[4 * pi + 3] = (f + 1) | ((n_slots + 1) << 16); ++n_slots; }
  }
  if (n_slots > 16) return true;
  lm->n_slots = std::max(1, n_slots);
  lm->n_bnd = static_cast<int>(bnd.size() / 4);
  std::vector<float> blob;
  auto put_f = [&](const float* src, size_t n) {
    const int off = static_cast<int>(blob.size());
    blob.insert(blob.end(), src, src + n);
    blob.resize(round4(static_cast<int>(blob.size())), 0.0f);
    return off;
  };
  auto put_i = [&](const int* src, size_t n) {
    const int off = static_cast<int>(blob.size());
    blob.resize(blob.size() + n);
    if (n) std::memcpy(blob.data() + off, src, n * sizeof(int));
    blob.resize(round4(static_cast<int>(blob.size())), 0.0f);
    return off;
  };
  lm->off_wtab = put_f(wtab.data(), wtab.size());
  lm->off_pieces = put_i(ptab.data(), ptab.size());
  lm->off_wrange = put_i(wrange.data(), wrange.size());
  if (bnd.empty()) bnd.assign(4, 0);
  lm->off_bnd = put_i(bnd.data(), bnd.size());
  lm->blob_f4 = static_cast<int>(blob.size() / 4);
  if (cudaMalloc(reinterpret_cast<void**>(&lm->blob_dev), blob.size() * sizeof(float)) != cudaSuccess) return false;
  if (cudaMemcpy(lm->blob_dev, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return false;
  lm->ok = 1;
  return true;
}

// ---- tables of the tensor-core path (tc_kernel.cu), n_fft = 512, int16 audio ---------------------------------
// Pass 1 of the 16 x 16 decomposition of the 256-point complex FFT behind the real 512-point FFT, per residue b:
//   Y_b[c] = W256^(b c) * sum_a W16^(a c) * (w[2n] x[2n] + i w[2n+1] x[2n+1]),  n = b + 16 a
// as a real 32 x 32 matrix M_b (rows 2c + re/im, columns 2a + even/odd sample) in float64, stored as float16 pairs:
// 16 M = MH1 + MH2 (value + residual); see the head of tc_kernel.cu for the operand split they multiply.
// The mel bank in segment form (as for the tiles path), dealt out to the 4 warps of a TMEM lane quarter; a filter whose
// terms all lie in one share is finished in registers ("direct"), the others go through boundary slots.
static bool build_tc_tables(asr_plan* pl, const std::vector<double>& mel_f, const std::vector<double>& fftfreqs) {
  const asr_mfcc_params& p = pl->prm;
  pl->tc_ok = 0;
  if (!pl->tl_ok || (p.hop_length % 32) != 0 || p.n_mels > 254) return true;
  const int n_bins = pl->n_bins, n_mels = p.n_mels, wl = pl->win_length, lpad = (p.n_fft - wl) / 2;
  auto wind = [&](int i) -> double {            // the float64 window librosa multiplies the frames with
    const int n = i - lpad;
    if (n < 0 || n >= wl) return 0.0;
    const double a0 = p.window == ASR_WIN_HANN ? 0.5 : 0.54, a1 = 1.0 - a0;
    return wl == 1 ? 1.0 : a0 - a1 * std::cos(2.0 * kPi * n / wl);
  };
  std::vector<__half> mats(static_cast<size_t>(16) * 2 * 32 * 32);
  auto put = [&](int b, int mat, int n, int k, double v) {
    mats[(((static_cast<size_t>(b) * 2 + mat) * 4 + k / 8) * 32 + n) * 8 + k % 8] = __float2half_rn(static_cast<float>(v));
  };
  for (int b = 0; b < 16; ++b)
    for (int c = 0; c < 16; ++c)
      for (int a = 0; a < 16; ++a) {
        const double th = 2.0 * kPi * (b * c / 256.0 + a * c / 16.0);
        const double tr = std::cos(th), ti = -std::sin(th);
        const int n = b + 16 * a;
        const double we = wind(2 * n), wo = wind(2 * n + 1);
        const double m[2][2] = {{tr * we, -ti * wo}, {ti * we, tr * wo}};      // [re/im of Y][even/odd sample]
        for (int ri = 0; ri < 2; ++ri)
          for (int eo = 0; eo < 2; ++eo) {
            const double mh = 16.0 * m[ri][eo];
            const __half h1 = __float2half_rn(static_cast<float>(mh));
            put(b, 0, 2 * c + ri, 2 * a + eo, mh);
            put(b, 1, 2 * c + ri, 2 * a + eo, mh - static_cast<double>(__half2float(h1)));
          }
      }
  if (cudaMalloc(&pl->tc_mats_dev, mats.size() * sizeof(__half)) != cudaSuccess) return false;
  if (cudaMemcpy(pl->tc_mats_dev, mats.data(), mats.size() * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess) return false;
  LaneMel lm;
  if (!build_lane_mel(pl, mel_f, fftfreqs, 1.0f / 1073741824.0f, &lm)) return false;   // the spectrum in TMEM is 2^30 |X|^2 (exact scaling)
  if (!lm.ok) return true;
  pl->tc_blob_dev = lm.blob_dev; pl->tc_blob_f4 = lm.blob_f4; pl->tc_off_wtab = lm.off_wtab; pl->tc_off_pieces = lm.off_pieces;
  pl->tc_off_wrange = lm.off_wrange; pl->tc_off_bnd = lm.off_bnd; pl->tc_n_bnd = lm.n_bnd; pl->tc_n_slots = lm.n_slots;
  if (tc_upload_constants() != cudaSuccess) return false;
  if (cudaHostAlloc(reinterpret_cast<void**>(&pl->tc_dbg_host), 16 * sizeof(int), cudaHostAllocMapped) == cudaSuccess) {
    std::memset(pl->tc_dbg_host, 0, 16 * sizeof(int));
    if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&pl->tc_dbg_dev), pl->tc_dbg_host, 0) != cudaSuccess) pl->tc_dbg_dev = nullptr;
  } else {
    cudaGetLastError();
    pl->tc_dbg_host = nullptr; pl->tc_dbg_dev = nullptr;
  }
  pl->tc_ok = 1;
  return true;
}

// ---- tables of the tensor-core dense DFT (tcdft_kernel.cu): FFT sizes without a register FFT ----------------
// B[r][n] = 16 w[n] cos(2 pi k n / n_fft) for r = k < bins, -16 w[n] sin(...) for r = nh + k; zero rows / columns as padding;
// B = B1 + B2 in float16; per K step of 16 samples one slab [B1 | B2][2 chunks of 8 samples][2 nh rows][8 halves].
static bool build_dft_tables(asr_plan* pl, const std::vector<double>& mel_f, const std::vector<double>& fftfreqs) {
  const asr_mfcc_params& p = pl->prm;
  pl->df_ok = 0;
  const int n_bins = pl->n_bins, nh = (n_bins + 15) & ~15, ks = (p.n_fft + 15) / 16;
  if (pl->fft_path || 2 * nh > 512 || nh > 256 || p.preemph != 0.0f || p.n_mfcc > 40) return true;
  const int wl = pl->win_length, lpad = (p.n_fft - wl) / 2;
  auto wind = [&](int i) -> double {
    const int n = i - lpad;
    if (n < 0 || n >= wl) return 0.0;
    const double a0 = p.window == ASR_WIN_HANN ? 0.5 : 0.54, a1 = 1.0 - a0;
    return wl == 1 ? 1.0 : a0 - a1 * std::cos(2.0 * kPi * n / wl);
  };
  const size_t slab_halves = static_cast<size_t>(2) * 2 * (2 * nh) * 8;
  std::vector<__half> mats(slab_halves * ks, __float2half_rn(0.0f));
  for (int st = 0; st < ks; ++st)
    for (int k16 = 0; k16 < 16; ++k16) {
      const int n = 16 * st + k16;
      if (n >= p.n_fft) continue;
      const double w = wind(n);
      for (int k = 0; k < n_bins; ++k) {
        // exact phase reduction: k n mod n_fft keeps the argument small
        const double ang = 2.0 * kPi * static_cast<double>((static_cast<long long>(k) * n) % p.n_fft) / p.n_fft;
        const double v[2] = {16.0 * w * std::cos(ang), -16.0 * w * std::sin(ang)};
        for (int ri = 0; ri < 2; ++ri) {
          const int r = ri * nh + k;
          const __half h1 = __float2half_rn(static_cast<float>(v[ri]));
          const __half h2 = __float2half_rn(static_cast<float>(v[ri] - static_cast<double>(__half2float(h1))));
          const size_t base = slab_halves * st + (static_cast<size_t>(k16 / 8) * (2 * nh) + r) * 8 + k16 % 8;
          mats[base] = h1;
          mats[base + slab_halves / 2] = h2;
        }
      }
    }
  LaneMel lm;
  if (!build_lane_mel(pl, mel_f, fftfreqs, 1.0f / 256.0f, &lm)) return false;      // the spectrum in TMEM is 256 |X|^2
  if (!lm.ok) return true;
  pl->df_blob_dev = lm.blob_dev; pl->df_blob_f4 = lm.blob_f4; pl->df_off_wtab = lm.off_wtab; pl->df_off_pieces = lm.off_pieces;
  pl->df_off_wrange = lm.off_wrange; pl->df_off_bnd = lm.off_bnd; pl->df_n_bnd = lm.n_bnd; pl->df_n_slots = lm.n_slots;
  if (cudaMalloc(&pl->df_mats_dev, mats.size() * sizeof(__half)) != cudaSuccess) return false;
  if (cudaMemcpy(pl->df_mats_dev, mats.data(), mats.size() * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess) return false;
  pl->df_nh = nh; pl->df_ksteps = ks; pl->df_bslab = static_cast<int>(slab_halves * sizeof(__half));
  if (!pl->tc_dbg_host && cudaHostAlloc(reinterpret_cast<void**>(&pl->tc_dbg_host), 16 * sizeof(int), cudaHostAllocMapped) == cudaSuccess) {
    std::memset(pl->tc_dbg_host, 0, 16 * sizeof(int));
    if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&pl->tc_dbg_dev), pl->tc_dbg_host, 0) != cudaSuccess) pl->tc_dbg_dev = nullptr;
  }
  cudaGetLastError();
  pl->df_ok = 1;
  return true;
}

struct FftShape { int M, G, P; };
static bool fft_shape(int n_fft, FftShape* s) {
  switch (n_fft) {
    case 512: *s = {256, 16, 16}; return true;
    case 1024: *s = {512, 16, 32}; return true;
    case 2048: *s = {1024, 32, 32}; return true;
    default: return false;
  }
}

static int validate(const asr_mfcc_params& p) {
  auto bad = [](const char* m) { set_error(std::string("asr_plan_create: ") + m); return ASR_ERR_INVALID; };
  if (p.sr <= 0) return bad("sr must be positive");
  if (p.n_fft < 8 || p.n_fft > 8192) return bad("n_fft must be in [8, 8192]");
  if (p.win_length < 0 || p.win_length > p.n_fft) return bad("win_length must be in [0, n_fft]");
  if (p.hop_length < 1) return bad("hop_length must be >= 1");
  if (p.window != ASR_WIN_HANN && p.window != ASR_WIN_HAMMING) return bad("window must be hann or hamming");
  if (p.pad_mode != ASR_PAD_REFLECT && p.pad_mode != ASR_PAD_CONSTANT) return bad("pad_mode must be reflect or constant");
  if (p.fftfreq_mode != ASR_FFTFREQ_LINSPACE && p.fftfreq_mode != ASR_FFTFREQ_RFFTFREQ) return bad("bad fftfreq_mode");
  if (p.n_mels < 1 || p.n_mels > 512) return bad("n_mels must be in [1, 512]");
  if (p.n_mfcc < 1 || p.n_mfcc > p.n_mels) return bad("n_mfcc must be in [1, n_mels]");
  if (p.fmin < 0 || (p.fmax > 0 && p.fmax <= p.fmin)) return bad("need 0 <= fmin < fmax");
  if (!(p.amin > 0)) return bad("amin must be positive");
  if (p.lifter < 0) return bad("lifter must be >= 0");
  if (p.delta_orders < 0 || p.delta_orders > 2) return bad("delta_orders must be 0, 1 or 2");
  if (p.delta_orders > 0 && (p.delta_width < 3 || (p.delta_width & 1) == 0 || p.delta_width > 63))
    return bad("delta_width must be odd, in [3, 63]");
  return ASR_OK;
}

}  // namespace asr

using namespace asr;
static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" int asr_version(void) { return ASR_B200_VERSION; }
extern "C" const char* asr_last_error(void) { return g_last_error.c_str(); }
extern "C" int asr_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" int asr_plan_create(const asr_mfcc_params* params, asr_plan** plan_out) {
  if (!params || !plan_out) { set_error("asr_plan_create: null pointer"); return ASR_ERR_INVALID; }
  *plan_out = nullptr;
  const int vr = validate(*params);
  if (vr != ASR_OK) return vr;
  asr_plan* pl = new (std::nothrow) asr_plan();
  if (!pl) { set_error("asr_plan_create: out of host memory"); return ASR_ERR_ALLOC; }
  const asr_mfcc_params& p = *params;
  pl->prm = p;
  pl->win_length = p.win_length > 0 ? p.win_length : p.n_fft;
  pl->pad = p.center ? p.n_fft / 2 : 0;
  pl->n_bins = 1 + p.n_fft / 2;
  FftShape fs{0, 0, 0};
  pl->fft_path = fft_shape(p.n_fft, &fs) ? 1 : 0;
  pl->fb = pl->fft_path ? kWarps * (32 / fs.G) : kWarps;
  const int s_off = pl->fft_path ? 0 : round4(p.n_fft);
  // FFT paths: two half-warps work on adjacent frame buffers; a stride of 16 (mod 32) words puts their
  // 64-byte rows in complementary bank halves
  auto stride16 = [](int x) { int v = (x + 31) / 32 * 32 + 16; return v - 32 >= x ? v - 32 : v; };
  pl->frame_stride = pl->fft_path ? stride16(2 * (fs.M + fs.G)) : pitch_odd4(s_off + round4(pl->n_bins + 16));
  pl->chunk_cap = ((pl->fb - 1) * p.hop_length + p.n_fft + 8 + 7) & ~7;   // + up to 7 samples of alignment shift
  pl->lm_pitch = pitch_odd4(p.n_mels);
  pl->dct_pitch = round4(p.n_mels);

  // ---- window: scipy.signal.get_window(name, win_length, fftbins=True), centre-padded to n_fft ----
  pl->h_window.assign(p.n_fft, 0.0f);
  {
    const int wl = pl->win_length, lpad = (p.n_fft - wl) / 2;
    const double a0 = p.window == ASR_WIN_HANN ? 0.5 : 0.54, a1 = 1.0 - a0;
    for (int n = 0; n < wl; ++n)
      pl->h_window[lpad + n] = static_cast<float>(wl == 1 ? 1.0 : a0 - a1 * std::cos(2.0 * kPi * n / wl));
  }
  // ---- mel bank: dense float32 (librosa) -> contiguous supports -> float4 tasks ----
  std::vector<double> mel_f, fftfreqs;
  pl->h_mel_dense = mel_dense(p, pl->n_bins, &mel_f, &fftfreqs);
  std::vector<MelTask> tasks;
  std::vector<float> melw;
  std::vector<int> ftasks(2 * p.n_mels, 0);
  for (int i = 0; i < p.n_mels; ++i) {
    const float* row = pl->h_mel_dense.data() + static_cast<size_t>(i) * pl->n_bins;
    int lo = -1, hi = -1;
    for (int k = 0; k < pl->n_bins; ++k)
      if (row[k] != 0.0f) { if (lo < 0) lo = k; hi = k + 1; }
    ftasks[2 * i] = static_cast<int>(tasks.size());
    if (lo >= 0) {
      const int a = lo & ~3;
      const int quads = (hi - a + 3) / 4;
      for (int q0 = 0; q0 < quads; q0 += kMelChunkQuads) {
        MelTask t;
        t.filter = i;
        t.k_start = a + 4 * q0;
        t.n_quads = std::min(kMelChunkQuads, quads - q0);
        t.w_off = static_cast<int>(melw.size() / 4);
        for (int k = t.k_start; k < t.k_start + 4 * kMelChunkQuads; ++k)   // always 4 quads, zero beyond the support
          melw.push_back((k < pl->n_bins && k < t.k_start + 4 * t.n_quads) ? row[k] : 0.0f);
        tasks.push_back(t);
      }
    }
    ftasks[2 * i + 1] = static_cast<int>(tasks.size()) - ftasks[2 * i];
  }
  pl->n_tasks = static_cast<int>(tasks.size());
  // Task order for the per-warp mel stage: lanes of a quarter-warp read S with 16-byte loads, so tasks that
  // sit next to each other should start in different 16-byte bank groups ((k_start/4) mod 8); deal them out
  // round-robin over the 8 residues.  Padded with empty tasks to a multiple of the streams per frame.
  const int spf = 32 / (pl->fb / kWarps);
  std::vector<std::vector<int>> bucket(8);
  for (int t = 0; t < pl->n_tasks; ++t) bucket[(tasks[t].k_start / 4) & 7].push_back(t);
  std::vector<int> order;
  for (size_t round = 0; order.size() < tasks.size(); ++round)
    for (int r = 0; r < 8; ++r)
      if (round < bucket[r].size()) order.push_back(bucket[r][round]);
  pl->n_tasks_padded = std::max(spf, (pl->n_tasks + spf - 1) / spf * spf);
  pl->part_pitch = round4(pl->n_tasks + 2);                    // slot n_tasks swallows the padding tasks, slot n_tasks+1 stays 0
  // per filter: up to 8 partial slots in summation order, padded with the always-zero slot
  int max_tasks = 0;
  for (int i = 0; i < p.n_mels; ++i) max_tasks = std::max(max_tasks, ftasks[2 * i + 1]);
  pl->fixed_slots = max_tasks <= 8 ? 1 : 0;
  std::vector<int> fslots(8 * static_cast<size_t>(p.n_mels), pl->n_tasks + 1);
  for (int i = 0; i < p.n_mels; ++i)
    for (int u = 0; u < std::min(8, ftasks[2 * i + 1]); ++u) fslots[8 * i + u] = ftasks[2 * i] + u;
  std::vector<int> task_tab(2 * static_cast<size_t>(pl->n_tasks_padded), 0);
  std::vector<float> melw_q(static_cast<size_t>(kMelChunkQuads) * pl->n_tasks_padded * 4, 0.0f);
  for (int tp = 0; tp < pl->n_tasks_padded; ++tp) {
    if (tp < pl->n_tasks) {
      const int t = order[tp];
      task_tab[2 * tp] = tasks[t].k_start;
      task_tab[2 * tp + 1] = t;
      for (int q = 0; q < kMelChunkQuads; ++q)
        for (int e = 0; e < 4; ++e)
          melw_q[(static_cast<size_t>(q) * pl->n_tasks_padded + tp) * 4 + e] = melw[(static_cast<size_t>(tasks[t].w_off) + q) * 4 + e];
    } else {
      task_tab[2 * tp] = 0;
      task_tab[2 * tp + 1] = pl->n_tasks;
    }
  }
  // ---- DCT-II ortho rows [0, n_mfcc) with the lifter folded in ----
  pl->h_dct.assign(static_cast<size_t>(p.n_mfcc) * p.n_mels, 0.0f);
  std::vector<float> dct_p(static_cast<size_t>(round4(p.n_mfcc)) * pl->dct_pitch, 0.0f);   // zero rows pad the last group of 4
  for (int c = 0; c < p.n_mfcc; ++c) {
    const double sc = c == 0 ? std::sqrt(1.0 / p.n_mels) : std::sqrt(2.0 / p.n_mels);
    const double lift = p.lifter > 0 ? 1.0 + (p.lifter / 2.0) * std::sin(kPi * (c + 1) / p.lifter) : 1.0;
    for (int n = 0; n < p.n_mels; ++n) {
      const float v = static_cast<float>(sc * std::cos(kPi * c * (2 * n + 1) / (2.0 * p.n_mels)) * lift);
      pl->h_dct[static_cast<size_t>(c) * p.n_mels + n] = v;
      dct_p[static_cast<size_t>(c) * pl->dct_pitch + n] = v;
    }
  }
  // ---- delta taps ----
  pl->h_taps.assign(static_cast<size_t>(std::max(1, p.delta_orders)) * std::max(1, p.delta_width), 0.0f);
  for (int o = 1; o <= p.delta_orders; ++o) {
    std::vector<double> t(p.delta_width);
    if (!savgol_taps(p.delta_width, o, t.data())) {
      set_error("asr_plan_create: singular Savitzky-Golay system");
      delete pl;
      return ASR_ERR_INVALID;
    }
    for (int j = 0; j < p.delta_width; ++j) pl->h_taps[static_cast<size_t>(o - 1) * p.delta_width + j] = static_cast<float>(t[j]);
  }
  // ---- FFT twiddles ----
  std::vector<float> twp, twu;
  if (pl->fft_path) {
    const int stride = 2 * fs.P + 4;
    twp.assign(static_cast<size_t>(fs.G) * stride, 0.0f);
    for (int n1 = 0; n1 < fs.G; ++n1)
      for (int k2 = 0; k2 < fs.P; ++k2) {
        const double ang = 2.0 * kPi * (static_cast<double>(n1) * k2) / fs.M;
        twp[static_cast<size_t>(n1) * stride + 2 * k2] = static_cast<float>(std::cos(ang));
        twp[static_cast<size_t>(n1) * stride + 2 * k2 + 1] = static_cast<float>(-std::sin(ang));
      }
    twu.assign(2 * (fs.M / 2 + 1), 0.0f);
    for (int k = 0; k <= fs.M / 2; ++k) {
      const double ang = 2.0 * kPi * k / p.n_fft;
      twu[2 * k] = static_cast<float>(-0.5 * std::sin(ang));
      twu[2 * k + 1] = static_cast<float>(-0.5 * std::cos(ang));
    }
  } else {
    twu.assign(2 * static_cast<size_t>(p.n_fft), 0.0f);
    for (int j = 0; j < p.n_fft; ++j) {
      const double ang = 2.0 * kPi * j / p.n_fft;
      twu[2 * j] = static_cast<float>(std::cos(ang));
      twu[2 * j + 1] = static_cast<float>(std::sin(ang));
    }
  }
  // ---- pack the blob (every section 16-byte aligned) ----
  std::vector<float> blob;
  auto put_f = [&](const float* src, size_t n) {
    const int off = static_cast<int>(blob.size());
    blob.insert(blob.end(), src, src + n);
    blob.resize(round4(static_cast<int>(blob.size())), 0.0f);
    return off;
  };
  auto put_i = [&](const int* src, size_t n) {
    const int off = static_cast<int>(blob.size());
    blob.resize(blob.size() + n);
    if (n) std::memcpy(blob.data() + off, src, n * sizeof(int));
    blob.resize(round4(static_cast<int>(blob.size())), 0.0f);
    return off;
  };
  pl->off_window = put_f(pl->h_window.data(), pl->h_window.size());
  {
    std::vector<float> wi(pl->h_window);
    for (float& v : wi) v *= (1.0f / 32768.0f);      // exact: int16 samples are scaled through the window
    pl->off_window_i16 = put_f(wi.data(), wi.size());
  }
  pl->off_twp = put_f(twp.data(), twp.size());
  pl->off_twu = put_f(twu.data(), twu.size());
  pl->off_tasks = put_i(task_tab.data(), task_tab.size());
  pl->off_melw = put_f(melw_q.data(), melw_q.size());
  pl->off_ftasks = put_i(ftasks.data(), ftasks.size());
  pl->off_fslots = put_i(fslots.data(), fslots.size());
  pl->off_dct = put_f(dct_p.data(), dct_p.size());
  pl->off_taps = put_f(pl->h_taps.data(), pl->h_taps.size());
  pl->blob_floats = static_cast<int>(blob.size());

  cudaError_t e = cudaGetDevice(&pl->device);
  if (e == cudaSuccess) e = mfcc_kernel_init();
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&pl->blob_dev), blob.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(pl->blob_dev, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (pl->blob_dev) cudaFree(pl->blob_dev);
    delete pl;
    return cuda_fail(e, "asr_plan_create (no usable CUDA device? there is no CPU fallback)");
  }
  {
    cudaDeviceProp prop;
    pl->sm_count = (cudaGetDeviceProperties(&prop, pl->device) == cudaSuccess) ? prop.multiProcessorCount : 148;
  }
  if (!build_cepstra_tables(pl) || !build_frames_tables(pl, mel_f, fftfreqs, twp, twu) || !build_tile_tables_all(pl, mel_f, fftfreqs, twp, twu) ||
      !build_tc_tables(pl, mel_f, fftfreqs) || !build_dft_tables(pl, mel_f, fftfreqs)) {
    const cudaError_t e2 = cudaGetLastError();
    asr_plan_destroy(pl);
    return cuda_fail(e2, "asr_plan_create (frames-path tables)");
  }
  *plan_out = pl;
  return ASR_OK;
}

void free_host_state(asr_plan* plan);   // host-buffer pipeline state (defined beside asr_mfcc_batch_host)

extern "C" void asr_plan_destroy(asr_plan* plan) {
  if (!plan) return;
  if (plan->blob_dev) cudaFree(plan->blob_dev);
  if (plan->fr_blob_dev) cudaFree(plan->fr_blob_dev);
  for (int c = 0; c < 2; ++c)
    if (plan->tl[c].blob_dev) cudaFree(plan->tl[c].blob_dev);
  if (plan->ws_dev) cudaFree(plan->ws_dev);
  if (plan->cep_dev) cudaFree(plan->cep_dev);
  if (plan->df_mats_dev) cudaFree(plan->df_mats_dev);
  if (plan->df_blob_dev) cudaFree(plan->df_blob_dev);
  if (plan->tc_mats_dev) cudaFree(plan->tc_mats_dev);
  if (plan->tc_blob_dev) cudaFree(plan->tc_blob_dev);
  if (plan->tc_dbg_host) cudaFreeHost(plan->tc_dbg_host);
  free_host_state(plan);
  delete plan;
}

extern "C" int32_t asr_plan_num_frames(const asr_plan* plan, int64_t length) {
  if (!plan || length < 0) return 0;
  if (plan->prm.pad_mode == ASR_PAD_REFLECT && plan->pad > 0 && length <= plan->pad) return 0;
  const int64_t padded = length + 2 * static_cast<int64_t>(plan->pad);
  if (padded < plan->prm.n_fft) return 0;
  return static_cast<int32_t>(1 + (padded - plan->prm.n_fft) / plan->prm.hop_length);
}

extern "C" int32_t asr_plan_feature_rows(const asr_plan* plan) {
  return plan ? plan->prm.n_mfcc * (1 + plan->prm.delta_orders) : 0;
}

extern "C" int32_t asr_plan_uses_fft(const asr_plan* plan) { return plan ? plan->fft_path : 0; }

extern "C" int asr_plan_get_tables(const asr_plan* plan, float* window, float* mel_dense_out, float* dct,
                                   float* delta_taps) {
  if (!plan) { set_error("asr_plan_get_tables: null plan"); return ASR_ERR_INVALID; }
  if (window) std::memcpy(window, plan->h_window.data(), plan->h_window.size() * sizeof(float));
  if (mel_dense_out) std::memcpy(mel_dense_out, plan->h_mel_dense.data(), plan->h_mel_dense.size() * sizeof(float));
  if (dct) std::memcpy(dct, plan->h_dct.data(), plan->h_dct.size() * sizeof(float));
  if (delta_taps && plan->prm.delta_orders > 0)
    std::memcpy(delta_taps, plan->h_taps.data(), sizeof(float) * plan->prm.delta_orders * plan->prm.delta_width);
  return ASR_OK;
}

// ---- block-pipelined path: shared-memory layout, workspace layout, launch ----
namespace {
struct FrLayout { int sm_aud, sm_S, sm_xb, sm_part, sm_raw, aud_cap, xb_stride, max_runs, smem_bytes, async_stage; };
// Shared-memory layout of the frames kernel.  Asynchronous staging (raw bytes of the next block copied in with cp.async:
// one sample buffer, half-size exchange buffers, a raw buffer) is taken when the alignment rules hold and it fits;
// otherwise the synchronous layout (two sample buffers filled through registers).
bool frames_layout(const asr_plan* plan, int dtype, int noise_mode, bool aligned, FrLayout* lo) {
  const asr_mfcc_params& p = plan->prm;
  const int hop = p.hop_length;
  const int esz = dtype == ASR_I16 ? 2 : (dtype == ASR_F32 ? 4 : 8);
  const bool can_async = aligned && noise_mode != ASR_NOISE_MIXTURE && p.preemph == 0.0f && (hop % 8) == 0 && (plan->pad % 8) == 0;
  for (int async_stage = can_async ? 1 : 0; async_stage >= 0; --async_stage) {
    for (int runs = kFrMaxRuns; runs >= (async_stage ? 2 : 1); --runs) {
      int off = 4 * plan->fr_blob_f4;
      lo->aud_cap = round4(kFrBlock * hop + runs * (512 - std::min(hop, 512)) + 4 * runs + 8);
      lo->sm_aud = off; off += (async_stage ? 1 : 2) * lo->aud_cap;
      lo->sm_S = off; off += 2 * kFrBlock * plan->fr_s_pitch;
      off = round4(off);
      lo->xb_stride = async_stage ? 272 : plan->frame_stride;      // 8 x 17 float2 (two-round exchange) : 16 x 17 float2 + bank split
      lo->sm_xb = off; off += kFrWarps * 2 * lo->xb_stride;
      lo->sm_part = off; off += 2 * plan->fr_n_refs * 33;
      off = round4(off);
      lo->sm_raw = off;
      if (async_stage) {
        // samples a block can need raw: 32 hops + per run the frame overlap, the reflection sources of a lone edge frame
        // and the alignment slack
        const long long samples = static_cast<long long>(kFrBlock) * hop + static_cast<long long>(runs) * (512 - std::min(hop, 512) + plan->pad + 1 + 16);
        const long long bytes = samples * (esz + (noise_mode != ASR_NOISE_NONE ? 8 : 0)) + 64LL * runs;
        off += static_cast<int>((bytes + 3) / 4);
      }
      lo->max_runs = runs;
      lo->async_stage = async_stage;
      lo->smem_bytes = 4 * off;
      if (lo->smem_bytes + 7168 <= kMaxSmemBytes) return true;      // + static shared memory (descriptor ring, clip cache)
    }
  }
  return false;
}
struct TlLayout { int sm_aud, sm_S, sm_part, sm_raw, max_runs, smem_bytes, nw, cfg, aud_cap; };
// Shared-memory layout of the tile kernel: tables | window | frame samples of one block | FR slots (exchange buffer /
// spectrum row) | mel partial rows | raw bytes of the block being copied in.  The 8-warp shape (two CTAs per SM) is taken
// when two such CTAs fit one SM, else the 16-warp shape.
bool tiles_layout(const asr_plan* plan, int dtype, int noise_mode, TlLayout* lo) {
  const asr_mfcc_params& p = plan->prm;
  const int hop = p.hop_length;
  const int esz = dtype == ASR_I16 ? 2 : (dtype == ASR_F32 ? 4 : 8);
  const char* force = std::getenv("ASR_B200_TILE_WARPS");          // experiments: 8 or 16
  for (int cfg = 0; cfg < 2; ++cfg) {                              // measured: the 16-warp shape is (slightly) faster when it fits
    const int nw = cfg == 1 ? 8 : 16, fr = 2 * nw;
    if (force && std::atoi(force) != nw) continue;
    const asr_plan::TileTables& tt = plan->tl[cfg];
    const long long budget = cfg == 1 ? (228 * 1024 - 2 * 1024) / 2 - 3072 : kMaxSmemBytes - 4096;   // minus static shared memory
    for (int runs = kTlMaxRuns; runs >= 2; --runs) {
      long long off = 4LL * tt.blob_f4 + 512;
      const long long aud_cap = round4(fr * hop + runs * (512 - std::min(hop, 512)) + 4 * runs + 8);
      lo->sm_aud = static_cast<int>(off); off += 2 * aud_cap;       // two sample buffers: block i is transformed while i+1 is converted
      lo->aud_cap = static_cast<int>(aud_cap);
      lo->sm_S = static_cast<int>(off); off += fr * kTlRS;
      lo->sm_part = static_cast<int>(off); off += tt.npart * fr;
      lo->sm_raw = static_cast<int>(off);
      // samples a block can need raw: FR hops + per run the frame overlap, the reflection sources of a lone edge frame
      // and the alignment slack
      const long long samples = static_cast<long long>(fr) * hop + static_cast<long long>(runs) * (512 - std::min(hop, 512) + plan->pad + 1 + 16);
      const long long bytes = samples * (esz + (noise_mode != ASR_NOISE_NONE ? 8 : 0)) + 64LL * runs;
      off += (bytes + 3) / 4;
      lo->max_runs = runs; lo->nw = nw; lo->cfg = cfg;
      if (4 * off <= budget) { lo->smem_bytes = static_cast<int>(4 * off); return true; }
    }
  }
  return false;
}
struct TcLayout { int sm_hl, sm_a, sm_b, sm_slots, hl_stride, hl_rows, smem_bytes; };
// Shared-memory layout of the tensor-core kernel: mel tables | staging array HL[16][stride] (8 bytes per sample pair) |
// operand ring (2 x (hi tile + lo tile) of 8 KB) | the 16 matrix sets (64 KB, resident) | boundary slots.
bool tc_layout(const asr_plan* plan, TcLayout* lo) {
  const asr_mfcc_params& p = plan->prm;
  long long off = 16LL * plan->tc_blob_f4;
  off = (off + 127) & ~127LL;
  const long long fixed = 2 * (2 * 8192) + 16 * 4096 + 512LL * plan->tc_n_slots + 1024;
  const long long budget = kMaxSmemBytes - tc_static_smem_bytes() - off - fixed;
  // rows a tile needs when every sub-block of 16 frames is one run (+ alignment), and a generous allowance for ragged batches
  const int typical = 8 * ((15 * p.hop_length + 512 + 31) / 32 + 1);
  int rows = static_cast<int>(std::min<long long>(budget / (16 * 8) - 16, 2LL * typical));
  if (rows < typical) return false;
  int stride = rows + 1;
  while ((stride & 15) != 1) ++stride;
  lo->hl_rows = rows; lo->hl_stride = stride;
  lo->sm_hl = static_cast<int>(off); off += 16LL * stride * 8;
  off = (off + 127) & ~127LL;
  lo->sm_a = static_cast<int>(off); off += 2 * (2 * 8192);
  lo->sm_b = static_cast<int>(off); off += 16 * 4096;
  lo->sm_slots = static_cast<int>(off); off += 512LL * plan->tc_n_slots;
  lo->smem_bytes = static_cast<int>(off);
  return off + tc_static_smem_bytes() <= kMaxSmemBytes;
}
// TC: the TILES conditions + int16 audio, hop a multiple of 32, no mixture noise; selected explicitly (ASR_PATH_TC) or by AUTO
bool tc_path_usable(const asr_plan* plan, int n_clips, int max_length, int dtype, int noise_mode, bool aligned, TcLayout* lo) {
  if (!plan->tc_ok || plan->path != ASR_PATH_TC) return false;
  if (dtype != ASR_I16 || noise_mode == ASR_NOISE_MIXTURE || !aligned) return false;
  const long long frames = static_cast<long long>(n_clips) * std::max(0, asr_plan_num_frames(plan, max_length));
  if (frames >= (1ll << 31) - 256) return false;
  return tc_layout(plan, lo);
}
struct DfLayout { int sm_stage, sm_slots, smem_bytes; };
// tensor-core dense DFT: the TC path of plans without a register FFT (clean input: no fused noise, no pre-emphasis)
bool tcdft_path_usable(const asr_plan* plan, int n_clips, int max_length, int noise_mode, DfLayout* lo) {
  if (!plan->df_ok || !(plan->path == ASR_PATH_TC || plan->path == ASR_PATH_AUTO) || noise_mode != ASR_NOISE_NONE) return false;
  const long long frames = static_cast<long long>(n_clips) * std::max(0, asr_plan_num_frames(plan, max_length));
  if (frames >= (1ll << 31) - 256) return false;
  long long off = 16LL * plan->df_blob_f4;
  off = (off + 127) & ~127LL;
  lo->sm_stage = static_cast<int>(off); off += static_cast<long long>(tcdft_stages()) * (2 * 4096 + plan->df_bslab);
  lo->sm_slots = static_cast<int>(off); off += 512LL * plan->df_n_slots;
  lo->smem_bytes = static_cast<int>(off);
  return off + tcdft_static_smem_bytes() <= kMaxSmemBytes;
}
struct WsLayout { size_t off_fstart, off_nframes, off_clipmax, off_lm, bytes; };
WsLayout ws_layout(const asr_plan* plan, int n_clips, int max_length) {
  WsLayout w;
  auto up = [](size_t x) { return (x + 255) & ~static_cast<size_t>(255); };
  const size_t frames = static_cast<size_t>(n_clips) * static_cast<size_t>(std::max(0, asr_plan_num_frames(plan, max_length)));
  w.off_fstart = 0;
  w.off_nframes = up(sizeof(int) * (static_cast<size_t>(n_clips) + 1));
  w.off_clipmax = w.off_nframes + up(sizeof(int) * static_cast<size_t>(n_clips));
  w.off_lm = w.off_clipmax + up(sizeof(float) * static_cast<size_t>(n_clips));
  w.bytes = w.off_lm + up(sizeof(float) * (frames + 32) * plan->fr_lm_pitch + 16);   // either layout: [frames][lm_pitch] or [n_mels][frames rounded up to 32]
  return w;
}
bool frames_path_usable(const asr_plan* plan, int n_clips, int max_length, int dtype, int noise_mode, bool aligned,
                        FrLayout* lo) {
  if (!plan->fr_ok || plan->path == ASR_PATH_CLIP) return false;
  const long long frames = static_cast<long long>(n_clips) * std::max(0, asr_plan_num_frames(plan, max_length));
  if (frames >= (1ll << 31) - 64) return false;                  // flattened frame index is int32
  return frames_layout(plan, dtype, noise_mode, aligned, lo);
}
// TILES: n_fft = 512, hop and pad multiples of 8, no pre-emphasis, no mixture noise, 16-byte aligned arrays
bool tiles_path_usable(const asr_plan* plan, int n_clips, int max_length, int dtype, int noise_mode, bool aligned,
                              TlLayout* lo) {
  if (!plan->tl_ok || !(plan->path == ASR_PATH_AUTO || plan->path == ASR_PATH_TILES || plan->path == ASR_PATH_TC)) return false;
  if (noise_mode == ASR_NOISE_MIXTURE || !aligned) return false;
  const long long frames = static_cast<long long>(n_clips) * std::max(0, asr_plan_num_frames(plan, max_length));
  if (frames >= (1ll << 31) - 64) return false;                  // flattened frame index is int32
  return tiles_layout(plan, dtype, noise_mode, lo);
}

// Per-clip kernel (mfcc_kernel.cu): dynamic shared memory of one CTA that keeps the log-mel rows of t_cap frames
long long clip_smem_bytes(const asr_plan* plan, int t_cap, bool want_cbuf) {
  const asr_mfcc_params& p = plan->prm;
  const int half = p.delta_orders > 0 ? p.delta_width / 2 : 0;
  long long off = plan->blob_floats;
  off += static_cast<long long>(plan->fb) * plan->frame_stride;
  off += static_cast<long long>(plan->fb) * plan->part_pitch;
  off += static_cast<long long>(t_cap) * plan->lm_pitch;
  off += want_cbuf ? static_cast<long long>(p.n_mfcc) * round4(t_cap + 2 * half + 1) : 0;
  off += 32;
  return off * 4;
}
// Clips whose log-mel matrix does not fit one CTA: chunks of frames per CTA + the cepstra kernels of the block-pipelined
// paths (through the log-mel workspace) instead of a thread-block cluster per clip.  `chunk` = frames per CTA (0: not used).
int clip_chunk_frames(const asr_plan* plan, int max_length) {
  const asr_mfcc_params& p = plan->prm;
  if (!plan->cep_dev || p.n_mfcc > 40) return 0;                          // no cepstra tables for the workspace kernels
  if (std::getenv("ASR_B200_CLIP_CLUSTERS")) return 0;                    // experiments: the cluster form
  const int t_all = std::max(1, asr_plan_num_frames(plan, max_length));
  if (clip_smem_bytes(plan, t_all, p.delta_orders > 0) <= kMaxSmemBytes) return 0;   // one CTA per clip is enough
  const int n_chunks = (t_all + 127) / 128;
  int fc = (t_all + n_chunks - 1) / n_chunks;
  fc = (fc + 15) & ~15;                                                    // whole warp iterations (8 warps x 2 frames)
  while (fc > 16 && clip_smem_bytes(plan, fc, false) > kMaxSmemBytes) fc -= 16;
  return clip_smem_bytes(plan, fc, false) <= kMaxSmemBytes ? fc : 0;
}

}  // namespace

extern "C" size_t asr_mfcc_workspace_bytes(const asr_plan* plan, int32_t n_clips, int32_t max_length) {
  FrLayout lo;
  if (!plan || n_clips <= 0 || max_length < 0) return 0;
  TlLayout tlo;
  TcLayout tcl;
  DfLayout dfl;
  if (tcdft_path_usable(plan, n_clips, max_length, ASR_NOISE_NONE, &dfl)) return ws_layout(plan, n_clips, max_length).bytes;
  if (!frames_path_usable(plan, n_clips, max_length, ASR_F64, ASR_NOISE_NONE, false, &lo) &&
      !tiles_path_usable(plan, n_clips, max_length, ASR_I16, ASR_NOISE_NONE, true, &tlo) &&
      !tc_path_usable(plan, n_clips, max_length, ASR_I16, ASR_NOISE_NONE, true, &tcl))
    return clip_chunk_frames(plan, max_length) > 0 ? ws_layout(plan, n_clips, max_length).bytes : 0;
  return ws_layout(plan, n_clips, max_length).bytes;
}

static bool frames_path_wanted(const asr_plan* plan, bool noisy) {
  return plan->path == ASR_PATH_FRAMES || plan->path == ASR_PATH_TILES || plan->path == ASR_PATH_TC || (plan->path == ASR_PATH_AUTO && noisy);
}
extern "C" int32_t asr_plan_launches(const asr_plan* plan, int32_t noisy) {
  FrLayout lo;
  TlLayout tlo;
  TcLayout tcl;
  DfLayout dfl;
  if (plan && tcdft_path_usable(plan, 1, 0, noisy ? ASR_NOISE_WHITE : ASR_NOISE_NONE, &dfl)) return 3;
  if (plan && tc_path_usable(plan, 1, 0, ASR_I16, noisy ? ASR_NOISE_WHITE : ASR_NOISE_NONE, true, &tcl)) return 3;
  if (plan && tiles_path_usable(plan, 1, 0, ASR_F64, noisy ? ASR_NOISE_WHITE : ASR_NOISE_NONE, true, &tlo)) return 3;
  return (plan && plan->fr_ok && plan->path != ASR_PATH_CLIP && frames_path_wanted(plan, noisy != 0) &&
          frames_layout(plan, ASR_F64, ASR_NOISE_NONE, false, &lo)) ? 3 : 1;
}

extern "C" int32_t asr_plan_path_used(const asr_plan* plan, int32_t dtype, int32_t noise_mode) {
  if (!plan) return ASR_PATH_CLIP;
  FrLayout lo;
  TlLayout tlo;
  TcLayout tcl;
  DfLayout dfl;
  if (tcdft_path_usable(plan, 1, 0, noise_mode, &dfl)) return ASR_PATH_TC;
  if (tc_path_usable(plan, 1, 0, dtype, noise_mode, true, &tcl)) return ASR_PATH_TC;
  if (tiles_path_usable(plan, 1, 0, dtype, noise_mode, true, &tlo)) return ASR_PATH_TILES;
  if (frames_path_wanted(plan, noise_mode != ASR_NOISE_NONE) && frames_path_usable(plan, 1, 0, dtype, noise_mode, true, &lo)) return ASR_PATH_FRAMES;
  return ASR_PATH_CLIP;
}

extern "C" int32_t asr_plan_debug_word(const asr_plan* plan, int32_t i) {
  if (!plan || !plan->tc_dbg_host || i < 0 || i >= 16) return 0;
  return reinterpret_cast<volatile int*>(plan->tc_dbg_host)[i];
}

extern "C" int asr_plan_set_stage_probe(asr_plan* plan, float* staged_dev) {
  if (!plan) { set_error("asr_plan_set_stage_probe: null plan"); return ASR_ERR_INVALID; }
  plan->stage_probe = staged_dev;
  return ASR_OK;
}

extern "C" int asr_plan_set_path(asr_plan* plan, int32_t path) {
  if (!plan || path < ASR_PATH_AUTO || path > ASR_PATH_TC) { set_error("asr_plan_set_path: bad argument"); return ASR_ERR_INVALID; }
  plan->path = path;
  return ASR_OK;
}

static int launch_common(const asr_plan* plan, const void* audio_dev, int32_t dtype, const int64_t* offsets_dev,
                         const int32_t* lengths_dev, int32_t n_clips, int32_t max_length, const asr_noise* noise,
                         void* out_dev, int32_t out_dtype, int32_t out_frames, int32_t* status_dev,
                         void* workspace_dev, size_t workspace_bytes, void* stream,
                         int logmel_only, const char* who) {
  auto bad = [&](const char* m) { set_error(std::string(who) + ": " + m); return ASR_ERR_INVALID; };
  if (!plan) return bad("null plan");
  if (n_clips < 0) return bad("negative n_clips");
  if (n_clips == 0) return ASR_OK;
  if (!audio_dev || !offsets_dev || !lengths_dev || !out_dev) return bad("null pointer");
  if (dtype < ASR_I16 || dtype > ASR_F64) return bad("dtype must be ASR_I16, ASR_F32 or ASR_F64");
  if (out_dtype != ASR_F32 && out_dtype != ASR_F64) return bad("out_dtype must be ASR_F32 or ASR_F64");
  if (out_frames < 1) return bad("out_frames must be >= 1");
  if (max_length < 0) return bad("negative max_length");
  const asr_mfcc_params& p = plan->prm;
  KParams kp;
  std::memset(&kp, 0, sizeof(kp));
  kp.audio = audio_dev;
  kp.offsets = reinterpret_cast<const long long*>(offsets_dev);
  kp.lengths = lengths_dev;
  kp.dtype = dtype;
  kp.n_clips = n_clips;
  kp.noise_mode = ASR_NOISE_NONE;
  if (noise && noise->mode != ASR_NOISE_NONE) {
    if (noise->mode == ASR_NOISE_WHITE) {
      if (!noise->z_dev || !noise->sigma_dev) return bad("white noise needs z_dev and sigma_dev");
    } else if (noise->mode == ASR_NOISE_MIXTURE) {
      if (!noise->z_dev || !noise->z2_dev) return bad("mixture noise needs z_dev (selector) and z2_dev (carrier)");
    } else {
      return bad("unknown noise mode");
    }
    kp.noise_mode = noise->mode;
    kp.z = noise->z_dev; kp.z2 = noise->z2_dev; kp.sigma = noise->sigma_dev;
    kp.mix_p = noise->p; kp.mix_s0 = noise->sigma0; kp.mix_s1 = noise->sigma1;
  }
  {
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    kp.vec_ok = al16(audio_dev) && (kp.noise_mode == ASR_NOISE_NONE || (al16(kp.z) && (kp.noise_mode != ASR_NOISE_MIXTURE || al16(kp.z2))));
  }
  // ---- FFT sizes without a register FFT (441): tensor-core dense DFT (frame prefix -> tcdft -> cepstra) ----
  DfLayout dfl;
  if (tcdft_path_usable(plan, n_clips, max_length, kp.noise_mode, &dfl)) {
    const WsLayout wl = ws_layout(plan, n_clips, max_length);
    char* ws = static_cast<char*>(workspace_dev);
    if (ws) {
      if (workspace_bytes < wl.bytes || (reinterpret_cast<uintptr_t>(ws) & 15)) return bad("workspace too small or not 16-byte aligned");
    } else {
      asr_plan* mp = const_cast<asr_plan*>(plan);                // plan-owned scratch (documented: no concurrent launches)
      if (mp->ws_bytes < wl.bytes) {
        if (mp->ws_dev) ASR_CUDA_TRY(cudaFree(mp->ws_dev));
        mp->ws_dev = nullptr; mp->ws_bytes = 0;
        ASR_CUDA_TRY(cudaMalloc(&mp->ws_dev, wl.bytes));
        mp->ws_bytes = wl.bytes;
      }
      ws = static_cast<char*>(mp->ws_dev);
    }
    FParams fp;
    std::memset(&fp, 0, sizeof(fp));
    fp.audio = audio_dev; fp.offsets = kp.offsets; fp.lengths = lengths_dev; fp.dtype = dtype; fp.n_clips = n_clips;
    fp.noise_mode = ASR_NOISE_NONE;
    fp.out = out_dev; fp.out_f64 = out_dtype == ASR_F64; fp.out_frames = out_frames;
    fp.out_rows = p.n_mfcc * (1 + p.delta_orders); fp.logmel_only = logmel_only; fp.status = status_dev;
    fp.n_fft = p.n_fft; fp.hop = p.hop_length; fp.pad = plan->pad; fp.pad_mode = p.pad_mode;
    fp.n_mels = p.n_mels; fp.n_mfcc = p.n_mfcc; fp.delta_orders = p.delta_orders; fp.delta_width = p.delta_width;
    fp.top_db = p.top_db; fp.amin = p.amin; fp.preemph = p.preemph;
    fp.blob = reinterpret_cast<const float4*>(plan->df_blob_dev); fp.blob_f4 = plan->df_blob_f4;
    fp.off_wtab = plan->df_off_wtab; fp.off_steps = plan->df_off_pieces; fp.off_wrange = plan->df_off_wrange;
    fp.tc_off_bnd = plan->df_off_bnd; fp.tc_n_bnd = plan->df_n_bnd;
    fp.tc_mats = plan->df_mats_dev; fp.tc_dbg = plan->tc_dbg_dev;
    fp.df_sm_stage = dfl.sm_stage; fp.tc_sm_slots = dfl.sm_slots;
    fp.df_nh = plan->df_nh; fp.df_ksteps = plan->df_ksteps; fp.df_bslab = plan->df_bslab;
    fp.fstart = reinterpret_cast<int*>(ws + wl.off_fstart);
    fp.nframes = reinterpret_cast<int*>(ws + wl.off_nframes);
    fp.clipmax = reinterpret_cast<float*>(ws + wl.off_clipmax);
    fp.lm = reinterpret_cast<float*>(ws + wl.off_lm);
    const long long frames = static_cast<long long>(n_clips) * std::max(0, asr_plan_num_frames(plan, max_length));
    fp.lm_stride = static_cast<int>((frames + 31) & ~31LL);
    fp.cep_blob = reinterpret_cast<const float4*>(plan->cep_dev);
    fp.cep_tab_f4 = plan->cep_tab_f4; fp.cep_off_cbuf = plan->cep_off_cbuf; fp.cep_off_taps = plan->cep_off_taps;
    fp.cep_off_col = plan->cep_smem_bytes / 4;
    fp.cep_small = (p.n_mels <= 32 && plan->h_dct_t.size() <= static_cast<size_t>(kCepSmallTab) && p.delta_orders == 0 && !logmel_only) ? 1 : 0;
    if (fp.cep_small) std::memcpy(fp.cep_dct, plan->h_dct_t.data(), plan->h_dct_t.size() * sizeof(float));
    { const char* dbg = std::getenv("ASR_B200_DBG_SKIP"); fp.dbg_skip = dbg ? std::atoi(dbg) : 0; }
    ASR_CUDA_TRY(launch_tcdft_path(fp, plan->sm_count, dfl.smem_bytes, plan->cep_smem_bytes + 4 * 128 * p.n_mels,
                                   std::max(1, asr_plan_num_frames(plan, max_length)), as_stream(stream)));
    return ASR_OK;
  }
  // ---- n_fft = 512, int16 audio: tensor-core path (frame prefix -> tc512 -> cepstra) ----
  TcLayout tcl;
  if (tc_path_usable(plan, n_clips, max_length, dtype, kp.noise_mode, kp.vec_ok != 0, &tcl)) {
    const WsLayout wl = ws_layout(plan, n_clips, max_length);
    char* ws = static_cast<char*>(workspace_dev);
    if (ws) {
      if (workspace_bytes < wl.bytes || (reinterpret_cast<uintptr_t>(ws) & 15)) return bad("workspace too small or not 16-byte aligned");
    } else {
      asr_plan* mp = const_cast<asr_plan*>(plan);                // plan-owned scratch (documented: no concurrent launches)
      if (mp->ws_bytes < wl.bytes) {
        if (mp->ws_dev) ASR_CUDA_TRY(cudaFree(mp->ws_dev));
        mp->ws_dev = nullptr; mp->ws_bytes = 0;
        ASR_CUDA_TRY(cudaMalloc(&mp->ws_dev, wl.bytes));
        mp->ws_bytes = wl.bytes;
      }
      ws = static_cast<char*>(mp->ws_dev);
    }
    FParams fp;
    std::memset(&fp, 0, sizeof(fp));
    fp.audio = audio_dev; fp.offsets = kp.offsets; fp.lengths = lengths_dev; fp.dtype = dtype; fp.n_clips = n_clips;
    fp.noise_mode = kp.noise_mode; fp.z = kp.z; fp.z2 = kp.z2; fp.sigma = kp.sigma;
    fp.out = out_dev; fp.out_f64 = out_dtype == ASR_F64; fp.out_frames = out_frames;
    fp.out_rows = p.n_mfcc * (1 + p.delta_orders); fp.logmel_only = logmel_only; fp.status = status_dev;
    fp.n_fft = p.n_fft; fp.hop = p.hop_length; fp.pad = plan->pad; fp.pad_mode = p.pad_mode;
    fp.n_mels = p.n_mels; fp.n_mfcc = p.n_mfcc; fp.delta_orders = p.delta_orders; fp.delta_width = p.delta_width;
    fp.top_db = p.top_db; fp.amin = p.amin; fp.preemph = p.preemph;
    fp.blob = reinterpret_cast<const float4*>(plan->tc_blob_dev); fp.blob_f4 = plan->tc_blob_f4;
    fp.off_wtab = plan->tc_off_wtab; fp.off_steps = plan->tc_off_pieces; fp.off_wrange = plan->tc_off_wrange;
    fp.tc_off_bnd = plan->tc_off_bnd; fp.tc_n_bnd = plan->tc_n_bnd;
    fp.tc_mats = plan->tc_mats_dev;
    fp.tc_dbg = plan->tc_dbg_dev;
    fp.tc_sm_hl = tcl.sm_hl; fp.tc_sm_a = tcl.sm_a; fp.tc_sm_b = tcl.sm_b; fp.tc_sm_slots = tcl.sm_slots;
    fp.tc_hl_stride = tcl.hl_stride; fp.tc_hl_rows = tcl.hl_rows;
    fp.fstart = reinterpret_cast<int*>(ws + wl.off_fstart);
    fp.nframes = reinterpret_cast<int*>(ws + wl.off_nframes);
    fp.clipmax = reinterpret_cast<float*>(ws + wl.off_clipmax);
    fp.lm = reinterpret_cast<float*>(ws + wl.off_lm);
    const long long frames = static_cast<long long>(n_clips) * std::max(0, asr_plan_num_frames(plan, max_length));
    fp.lm_stride = static_cast<int>((frames + 31) & ~31LL);
    fp.cep_blob = reinterpret_cast<const float4*>(plan->cep_dev);
    fp.cep_tab_f4 = plan->cep_tab_f4; fp.cep_off_cbuf = plan->cep_off_cbuf; fp.cep_off_taps = plan->cep_off_taps;
    fp.cep_off_col = plan->cep_smem_bytes / 4;
    fp.cep_small = (p.n_mels <= 32 && plan->h_dct_t.size() <= static_cast<size_t>(kCepSmallTab) && p.delta_orders == 0 && !logmel_only) ? 1 : 0;
    if (fp.cep_small) std::memcpy(fp.cep_dct, plan->h_dct_t.data(), plan->h_dct_t.size() * sizeof(float));
    { const char* dbg = std::getenv("ASR_B200_DBG_SKIP"); fp.dbg_skip = dbg ? std::atoi(dbg) : 0; }
    ASR_CUDA_TRY(launch_tc_path(fp, plan->sm_count, tcl.smem_bytes, plan->cep_smem_bytes + 4 * 128 * p.n_mels,
                                std::max(1, asr_plan_num_frames(plan, max_length)), as_stream(stream)));
    return ASR_OK;
  }
  // ---- n_fft = 512: TMA-staged block pipeline (frame prefix -> tiles -> cepstra) ----
  TlLayout tlo;
  if (tiles_path_usable(plan, n_clips, max_length, dtype, kp.noise_mode, kp.vec_ok != 0, &tlo)) {
    const WsLayout wl = ws_layout(plan, n_clips, max_length);
    char* ws = static_cast<char*>(workspace_dev);
    if (ws) {
      if (workspace_bytes < wl.bytes || (reinterpret_cast<uintptr_t>(ws) & 15)) return bad("workspace too small or not 16-byte aligned");
    } else {
      asr_plan* mp = const_cast<asr_plan*>(plan);                // plan-owned scratch (documented: no concurrent launches)
      if (mp->ws_bytes < wl.bytes) {
        if (mp->ws_dev) ASR_CUDA_TRY(cudaFree(mp->ws_dev));
        mp->ws_dev = nullptr; mp->ws_bytes = 0;
        ASR_CUDA_TRY(cudaMalloc(&mp->ws_dev, wl.bytes));
        mp->ws_bytes = wl.bytes;
      }
      ws = static_cast<char*>(mp->ws_dev);
    }
    FParams fp;
    std::memset(&fp, 0, sizeof(fp));
    fp.audio = audio_dev; fp.offsets = kp.offsets; fp.lengths = lengths_dev; fp.dtype = dtype; fp.n_clips = n_clips;
    fp.noise_mode = kp.noise_mode; fp.z = kp.z; fp.z2 = kp.z2; fp.sigma = kp.sigma;
    fp.out = out_dev; fp.out_f64 = out_dtype == ASR_F64; fp.out_frames = out_frames;
    fp.out_rows = p.n_mfcc * (1 + p.delta_orders); fp.logmel_only = logmel_only; fp.status = status_dev;
    fp.n_fft = p.n_fft; fp.hop = p.hop_length; fp.pad = plan->pad; fp.pad_mode = p.pad_mode;
    fp.n_mels = p.n_mels; fp.n_mfcc = p.n_mfcc; fp.delta_orders = p.delta_orders; fp.delta_width = p.delta_width;
    fp.top_db = p.top_db; fp.amin = p.amin; fp.preemph = p.preemph;
    const asr_plan::TileTables& tt = plan->tl[tlo.cfg];
    fp.blob = reinterpret_cast<const float4*>(tt.blob_dev); fp.blob_f4 = tt.blob_f4;
    fp.off_window = dtype == ASR_I16 ? tt.off_window_i16 : tt.off_window;   // int16 is staged unscaled, 2^-15 in the window
    fp.off_twp = tt.off_twp; fp.off_twu = tt.off_twu; fp.off_wtab = tt.off_wtab;
    fp.off_steps = tt.off_steps; fp.off_wrange = tt.off_wrange;
    fp.sm_aud = tlo.sm_aud; fp.sm_S = tlo.sm_S; fp.sm_part = tlo.sm_part; fp.sm_raw = tlo.sm_raw;
    fp.max_runs = tlo.max_runs; fp.vec_ok = 1; fp.aud_cap = tlo.aud_cap;
    fp.t_npart = tt.npart; fp.t_npc = tt.npc; fp.t_nw = tlo.nw; fp.t_smem_bytes = tlo.smem_bytes;
    fp.fstart = reinterpret_cast<int*>(ws + wl.off_fstart);
    fp.nframes = reinterpret_cast<int*>(ws + wl.off_nframes);
    fp.clipmax = reinterpret_cast<float*>(ws + wl.off_clipmax);
    fp.lm = reinterpret_cast<float*>(ws + wl.off_lm);
    const long long frames = static_cast<long long>(n_clips) * std::max(0, asr_plan_num_frames(plan, max_length));
    fp.lm_stride = static_cast<int>((frames + 31) & ~31LL);
    fp.cep_blob = reinterpret_cast<const float4*>(plan->cep_dev);
    fp.cep_tab_f4 = plan->cep_tab_f4; fp.cep_off_cbuf = plan->cep_off_cbuf; fp.cep_off_taps = plan->cep_off_taps;
    fp.cep_off_col = plan->cep_smem_bytes / 4;
    fp.cep_small = (p.n_mels <= 32 && plan->h_dct_t.size() <= static_cast<size_t>(kCepSmallTab) && p.delta_orders == 0 && !logmel_only) ? 1 : 0;
    if (fp.cep_small) std::memcpy(fp.cep_dct, plan->h_dct_t.data(), plan->h_dct_t.size() * sizeof(float));
    fp.stage_probe = plan->stage_probe;
    ASR_CUDA_TRY(launch_tiles_path(fp, plan->sm_count, tlo.smem_bytes, plan->cep_smem_bytes + 4 * 128 * p.n_mels,
                                   std::max(1, asr_plan_num_frames(plan, max_length)), as_stream(stream)));
    return ASR_OK;
  }
  // ---- n_fft = 512: block-pipelined path (frame prefix -> frames -> cepstra) ----
  FrLayout flo;
  if (frames_path_wanted(plan, kp.noise_mode != ASR_NOISE_NONE) &&
      frames_path_usable(plan, n_clips, max_length, dtype, kp.noise_mode, kp.vec_ok != 0, &flo)) {
    const WsLayout wl = ws_layout(plan, n_clips, max_length);
    char* ws = static_cast<char*>(workspace_dev);
    if (ws) {
      if (workspace_bytes < wl.bytes || (reinterpret_cast<uintptr_t>(ws) & 15)) return bad("workspace too small or not 16-byte aligned");
    } else {
      asr_plan* mp = const_cast<asr_plan*>(plan);                // plan-owned scratch (documented: no concurrent launches)
      if (mp->ws_bytes < wl.bytes) {
        if (mp->ws_dev) ASR_CUDA_TRY(cudaFree(mp->ws_dev));
        mp->ws_dev = nullptr; mp->ws_bytes = 0;
        ASR_CUDA_TRY(cudaMalloc(&mp->ws_dev, wl.bytes));
        mp->ws_bytes = wl.bytes;
      }
      ws = static_cast<char*>(mp->ws_dev);
    }
    FParams fp;
    std::memset(&fp, 0, sizeof(fp));
    fp.audio = audio_dev; fp.offsets = kp.offsets; fp.lengths = lengths_dev; fp.dtype = dtype; fp.n_clips = n_clips;
    fp.noise_mode = kp.noise_mode; fp.z = kp.z; fp.z2 = kp.z2; fp.sigma = kp.sigma;
    fp.mix_p = kp.mix_p; fp.mix_s0 = kp.mix_s0; fp.mix_s1 = kp.mix_s1;
    fp.out = out_dev; fp.out_f64 = out_dtype == ASR_F64; fp.out_frames = out_frames;
    fp.out_rows = p.n_mfcc * (1 + p.delta_orders); fp.logmel_only = logmel_only; fp.status = status_dev;
    fp.n_fft = p.n_fft; fp.hop = p.hop_length; fp.pad = plan->pad; fp.pad_mode = p.pad_mode;
    fp.n_mels = p.n_mels; fp.n_mfcc = p.n_mfcc; fp.delta_orders = p.delta_orders; fp.delta_width = p.delta_width;
    fp.top_db = p.top_db; fp.amin = p.amin; fp.preemph = p.preemph;
    fp.blob = reinterpret_cast<const float4*>(plan->fr_blob_dev); fp.blob_f4 = plan->fr_blob_f4;
    const bool unscaled_i16 = dtype == ASR_I16 && kp.noise_mode == ASR_NOISE_NONE;   // staged unscaled, 2^-15 in the window
    fp.off_window = unscaled_i16 ? plan->fr_off_window_i16 : plan->fr_off_window;
    fp.off_twp = plan->fr_off_twp; fp.off_twu = plan->fr_off_twu; fp.off_wtab = plan->fr_off_wtab;
    fp.off_pieces = plan->fr_off_pieces; fp.off_wrange = plan->fr_off_wrange; fp.off_frange = plan->fr_off_frange;
    fp.off_refs = plan->fr_off_refs;
    fp.sm_aud = flo.sm_aud; fp.sm_S = flo.sm_S; fp.sm_xb = flo.sm_xb; fp.sm_part = flo.sm_part; fp.sm_raw = flo.sm_raw;
    fp.async_stage = flo.async_stage;
    fp.aud_cap = flo.aud_cap; fp.s_pitch = plan->fr_s_pitch; fp.xb_stride = flo.xb_stride;
    fp.n_refs = plan->fr_n_refs; fp.max_runs = flo.max_runs;
    fp.vec_ok = kp.vec_ok && p.preemph == 0.0f && (p.hop_length % 4) == 0 && (plan->pad % 4) == 0 &&
                (!flo.async_stage || kp.noise_mode != ASR_NOISE_MIXTURE);
    fp.fstart = reinterpret_cast<int*>(ws + wl.off_fstart);
    fp.nframes = reinterpret_cast<int*>(ws + wl.off_nframes);
    fp.clipmax = reinterpret_cast<float*>(ws + wl.off_clipmax);
    fp.lm = reinterpret_cast<float*>(ws + wl.off_lm);
    fp.lm_pitch = plan->fr_lm_pitch;
    fp.cep_blob_f4 = plan->cep_blob_f4; fp.cep_tab_f4 = plan->cep_tab_f4;
    fp.cep_off_cbuf = plan->cep_off_cbuf; fp.cep_off_taps = plan->cep_off_taps;
    { const char* dbg = std::getenv("ASR_B200_DBG_SKIP"); fp.dbg_skip = dbg ? std::atoi(dbg) : 0; }
    ASR_CUDA_TRY(launch_frames_path(fp, plan->sm_count, flo.smem_bytes, plan->cep_smem_bytes,
                                    std::max(1, asr_plan_num_frames(plan, max_length)), as_stream(stream)));
    return ASR_OK;
  }
  kp.out = out_dev;
  kp.out_f64 = out_dtype == ASR_F64;
  kp.out_frames = out_frames;
  kp.out_rows = p.n_mfcc * (1 + p.delta_orders);
  kp.logmel_only = logmel_only;
  kp.status = status_dev;
  kp.n_fft = p.n_fft; kp.hop = p.hop_length; kp.pad = plan->pad; kp.pad_mode = p.pad_mode;
  kp.n_bins = plan->n_bins; kp.n_mels = p.n_mels; kp.n_mfcc = p.n_mfcc;
  kp.delta_orders = p.delta_orders; kp.delta_width = p.delta_width;
  kp.top_db = p.top_db; kp.amin = p.amin; kp.preemph = p.preemph;
  kp.lm_pitch = plan->lm_pitch; kp.dct_pitch = plan->dct_pitch;
  kp.fft_path = plan->fft_path; kp.fb = plan->fb; kp.frame_stride = plan->frame_stride; kp.chunk_cap = plan->chunk_cap;
  kp.n_tasks = plan->n_tasks; kp.n_tasks_padded = plan->n_tasks_padded; kp.part_pitch = plan->part_pitch;
  kp.blob = reinterpret_cast<const float4*>(plan->blob_dev);
  kp.blob_f4 = plan->blob_floats / 4;
  kp.off_window = plan->off_window; kp.off_window_i16 = plan->off_window_i16; kp.off_twp = plan->off_twp; kp.off_twu = plan->off_twu;
  kp.off_tasks = plan->off_tasks; kp.off_melw = plan->off_melw;
  kp.off_ftasks = plan->off_ftasks; kp.off_fslots = plan->off_fslots; kp.fixed_slots = plan->fixed_slots;
  kp.off_dct = plan->off_dct;
  kp.off_taps = plan->off_taps;
  // ---- cluster size + dynamic shared memory layout ----
  // One CTA holds the log-mel rows of ceil(T/cs) frames; grow the cluster until that fits, then keep
  // growing (up to 8) while the launch would leave most of the 148 SMs without a CTA.
  const int t_all = std::max(1, asr_plan_num_frames(plan, max_length));
  const int half = p.delta_orders > 0 ? p.delta_width / 2 : 0;
  const bool want_cbuf = p.delta_orders > 0 && !logmel_only;
  auto layout = [&](int cs, KParams& k) -> long long {
    const int t_cap = (t_all + cs - 1) / cs;
    int off = plan->blob_floats;
    k.sm_frames = off; off += plan->fb * plan->frame_stride;
    k.sm_part = off; off += plan->fb * plan->part_pitch;
    k.sm_lm = off; off += t_cap * plan->lm_pitch;
    k.cbuf_pitch = round4(t_cap + 2 * half + 1);
    k.sm_cbuf = off; off += want_cbuf ? p.n_mfcc * k.cbuf_pitch : 0;
    k.sm_red = off; off += 32;
    k.t_cap = t_cap;
    k.cluster_size = cs;
    return static_cast<long long>(off) * 4;
  };
  // ---- long clips: chunks of frames per CTA, log-mel rows through the workspace, cepstra kernels of the pipelined paths ----
  if (const int fc = clip_chunk_frames(plan, max_length)) {
    const long long frames = static_cast<long long>(n_clips) * t_all;
    if (frames < (1ll << 31) - 64) {
      const WsLayout wl = ws_layout(plan, n_clips, max_length);
      char* ws = static_cast<char*>(workspace_dev);
      if (ws) {
        if (workspace_bytes < wl.bytes || (reinterpret_cast<uintptr_t>(ws) & 15)) return bad("workspace too small or not 16-byte aligned");
      } else {
        asr_plan* mp = const_cast<asr_plan*>(plan);                // plan-owned scratch (documented: no concurrent launches)
        if (mp->ws_bytes < wl.bytes) {
          if (mp->ws_dev) ASR_CUDA_TRY(cudaFree(mp->ws_dev));
          mp->ws_dev = nullptr; mp->ws_bytes = 0;
          ASR_CUDA_TRY(cudaMalloc(&mp->ws_dev, wl.bytes));
          mp->ws_bytes = wl.bytes;
        }
        ws = static_cast<char*>(mp->ws_dev);
      }
      FParams fp;
      std::memset(&fp, 0, sizeof(fp));
      fp.audio = audio_dev; fp.offsets = kp.offsets; fp.lengths = lengths_dev; fp.dtype = dtype; fp.n_clips = n_clips;
      fp.noise_mode = ASR_NOISE_NONE;
      fp.out = out_dev; fp.out_f64 = out_dtype == ASR_F64; fp.out_frames = out_frames;
      fp.out_rows = p.n_mfcc * (1 + p.delta_orders); fp.logmel_only = logmel_only; fp.status = status_dev;
      fp.n_fft = p.n_fft; fp.hop = p.hop_length; fp.pad = plan->pad; fp.pad_mode = p.pad_mode;
      fp.n_mels = p.n_mels; fp.n_mfcc = p.n_mfcc; fp.delta_orders = p.delta_orders; fp.delta_width = p.delta_width;
      fp.top_db = p.top_db; fp.amin = p.amin; fp.preemph = p.preemph;
      fp.fstart = reinterpret_cast<int*>(ws + wl.off_fstart);
      fp.nframes = reinterpret_cast<int*>(ws + wl.off_nframes);
      fp.clipmax = reinterpret_cast<float*>(ws + wl.off_clipmax);
      fp.lm = reinterpret_cast<float*>(ws + wl.off_lm);
      fp.lm_stride = static_cast<int>((frames + 31) & ~31LL);
      fp.cep_blob = reinterpret_cast<const float4*>(plan->cep_dev);
      fp.cep_tab_f4 = plan->cep_tab_f4; fp.cep_off_cbuf = plan->cep_off_cbuf; fp.cep_off_taps = plan->cep_off_taps;
      fp.cep_off_col = plan->cep_smem_bytes / 4;
      fp.cep_small = (p.n_mels <= 32 && plan->h_dct_t.size() <= static_cast<size_t>(kCepSmallTab) && p.delta_orders == 0 && !logmel_only) ? 1 : 0;
      if (fp.cep_small) std::memcpy(fp.cep_dct, plan->h_dct_t.data(), plan->h_dct_t.size() * sizeof(float));
      layout(1, kp);                                               // offsets of the fixed parts
      kp.t_cap = fc; kp.cluster_size = 1;
      kp.sm_cbuf = kp.sm_lm + fc * plan->lm_pitch; kp.sm_red = kp.sm_cbuf;       // (no cepstra buffer in this mode)
      kp.chunk_mode = 1; kp.n_chunks = (t_all + fc - 1) / fc;
      kp.lm_global = fp.lm; kp.lm_stride = fp.lm_stride; kp.fstart = fp.fstart;
      const long long chunk_smem = clip_smem_bytes(plan, fc, false);
      ASR_CUDA_TRY(launch_frame_prefix(fp, as_stream(stream)));
      ASR_CUDA_TRY(launch_mfcc(kp, static_cast<int>(chunk_smem), as_stream(stream)));
      ASR_CUDA_TRY(launch_cepstra_tail(fp, plan->cep_smem_bytes + 4 * 128 * p.n_mels, t_all, as_stream(stream)));
      return ASR_OK;
    }
  }
  int cs = 1;
  long long smem_bytes = layout(cs, kp);
  while (smem_bytes > kMaxSmemBytes && cs < 16) { cs *= 2; smem_bytes = layout(cs, kp); }
  if (smem_bytes > kMaxSmemBytes) {
    set_error(std::string(who) + ": clip of " + std::to_string(max_length) + " samples (" + std::to_string(t_all) +
              " frames) needs " + std::to_string(smem_bytes) + " bytes of shared memory per CTA even with a 16-CTA"
              " cluster; limit is " + std::to_string(kMaxSmemBytes));
    return ASR_ERR_TOO_LARGE;
  }
  while (cs < 8 && static_cast<long long>(n_clips) * cs < 4 * 148 && (t_all + 2 * cs - 1) / (2 * cs) >= 2 * plan->fb) {
    cs *= 2;
    smem_bytes = layout(cs, kp);
  }
  {
    const cudaError_t le = launch_mfcc(kp, static_cast<int>(smem_bytes), as_stream(stream));
    if (le == cudaErrorLaunchOutOfResources && cs > 1) {
      cudaGetLastError();
      set_error(std::string(who) + ": a cluster of " + std::to_string(cs) + " CTAs with " + std::to_string(smem_bytes) +
                " bytes of shared memory each cannot be co-scheduled on this device (clip of " + std::to_string(max_length) +
                " samples, " + std::to_string(t_all) + " frames)");
      return ASR_ERR_TOO_LARGE;
    }
    ASR_CUDA_TRY(le);
  }
  return ASR_OK;
}

extern "C" int asr_mfcc_batch(const asr_plan* plan, const void* audio_dev, int32_t dtype, const int64_t* offsets_dev,
                              const int32_t* lengths_dev, int32_t n_clips, int32_t max_length, const asr_noise* noise,
                              void* out_dev, int32_t out_dtype, int32_t out_frames, int32_t* status_dev,
                              void* workspace_dev, size_t workspace_bytes, void* stream) {
  return launch_common(plan, audio_dev, dtype, offsets_dev, lengths_dev, n_clips, max_length, noise, out_dev,
                       out_dtype, out_frames, status_dev, workspace_dev, workspace_bytes, stream, 0, "asr_mfcc_batch");
}

extern "C" int asr_logmel_batch(const asr_plan* plan, const void* audio_dev, int32_t dtype, const int64_t* offsets_dev,
                                const int32_t* lengths_dev, int32_t n_clips, int32_t max_length, const asr_noise* noise,
                                float* out_dev, int32_t out_frames, int32_t* status_dev,
                                void* workspace_dev, size_t workspace_bytes, void* stream) {
  return launch_common(plan, audio_dev, dtype, offsets_dev, lengths_dev, n_clips, max_length, noise, out_dev, ASR_F32,
                       out_frames, status_dev, workspace_dev, workspace_bytes, stream, 1, "asr_logmel_batch");
}

// ------------------------------------------------------------------------------------------------
// The one collective of the sharded path behind the C-ABI: ncclAllGather of the standardisation messages, resolved from the
// NCCL library the process has already loaded (no link-time dependency, no second copy of NCCL).
namespace {
typedef int (*nccl_allgather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
nccl_allgather_fn resolve_nccl_allgather() {
  static nccl_allgather_fn fn = nullptr;
  static std::mutex mu;
  std::lock_guard<std::mutex> g(mu);
  if (fn) return fn;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    void* h = dlopen(n, RTLD_NOW | RTLD_NOLOAD);               // only a copy that is already in the process
    if (h) {
      fn = reinterpret_cast<nccl_allgather_fn>(dlsym(h, "ncclAllGather"));
      if (fn) return fn;
    }
  }
  fn = reinterpret_cast<nccl_allgather_fn>(dlsym(RTLD_DEFAULT, "ncclAllGather"));   // statically linked into the host program
  return fn;
}
}  // namespace

extern "C" int asr_cmvn_exchange_nccl(void* nccl_comm, const double* msg_dev, double* msgs_dev, int32_t n_cols, void* stream) {
  if (!nccl_comm || !msg_dev || !msgs_dev || n_cols < 1) { set_error("asr_cmvn_exchange_nccl: bad argument"); return ASR_ERR_INVALID; }
  const nccl_allgather_fn fn = resolve_nccl_allgather();
  if (!fn) { set_error("asr_cmvn_exchange_nccl: no NCCL library is loaded in this process (load the one that created nccl_comm first)"); return ASR_ERR_INVALID; }
  const int rc = fn(msg_dev, msgs_dev, static_cast<size_t>(3) * n_cols + 1, 8 /* ncclFloat64 */, nccl_comm, reinterpret_cast<cudaStream_t>(stream));
  if (rc != 0) { set_error("asr_cmvn_exchange_nccl: ncclAllGather failed with ncclResult_t " + std::to_string(rc)); return ASR_ERR_CUDA; }
  return ASR_OK;
}

// ------------------------------------------------------------------------------------------------
// Host-buffer pipeline: clips are cut into chunks; chunk i+1's H2D copy overlaps chunk i's kernels
// and chunk i-1's D2H copy (two streams, two sets of device buffers).  The streams, the device buffers and the pinned
// per-chunk descriptors live in the plan (grown on demand, freed by asr_plan_destroy), so a call allocates nothing once
// the plan has seen its batch shape; calls on one plan are serialised by a mutex (the plan's tables stay immutable).
namespace {
struct Slot {
  cudaStream_t st = nullptr;
  void* audio = nullptr;      size_t audio_cap = 0;
  long long* offsets = nullptr; int* lengths = nullptr; int* status = nullptr; size_t clip_cap = 0;
  void* out = nullptr;        size_t out_cap = 0;
  float* power = nullptr;     double* sigma = nullptr; size_t snr_cap = 0;
  double* z = nullptr;        size_t z_cap = 0;
  void* ws = nullptr;         size_t ws_cap = 0;
  long long* reb_pinned = nullptr; size_t reb_cap = 0;     // rebased offsets of the chunk in flight (pinned: no sync after the copy)
  int* len_pinned = nullptr;                               // its lengths (the caller's array may be pageable: that copy would block)
  cudaEvent_t uploaded = nullptr;                          // the chunk's descriptors have left reb_pinned
  template <typename T>
  static cudaError_t grow(T** p, size_t* cap, size_t need) {
    if (need <= *cap) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    const cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), need);
    if (e == cudaSuccess) *cap = need;
    return e;
  }
  void release() {
    if (ws) cudaFree(ws);
    if (audio) cudaFree(audio);
    if (offsets) cudaFree(offsets);
    if (lengths) cudaFree(lengths);
    if (out) cudaFree(out);
    if (status) cudaFree(status);
    if (power) cudaFree(power);
    if (sigma) cudaFree(sigma);
    if (z) cudaFree(z);
    if (reb_pinned) cudaFreeHost(reb_pinned);
    if (len_pinned) cudaFreeHost(len_pinned);
    if (uploaded) cudaEventDestroy(uploaded);
    if (st) cudaStreamDestroy(st);
  }
};
struct HostState {
  std::mutex mu;
  Slot slots[2];
  int device = -1;
  // per-clip status of the whole call, pinned: a device-to-host copy into the caller's (pageable) array would block the
  // host until the chunk has finished, i.e. serialise upload / kernels / download of consecutive chunks
  int* status_pinned = nullptr; size_t status_cap = 0;
};
}  // namespace

void free_host_state(asr_plan* plan) {
  if (!plan || !plan->host_state) return;
  HostState* hs = static_cast<HostState*>(plan->host_state);
  for (int s = 0; s < 2; ++s) hs->slots[s].release();
  if (hs->status_pinned) cudaFreeHost(hs->status_pinned);
  delete hs;
  plan->host_state = nullptr;
}

extern "C" int asr_mfcc_batch_host(const asr_plan* plan, const void* audio_host, int32_t dtype,
                                   const int64_t* offsets_host, const int32_t* lengths_host, int32_t n_clips,
                                   int32_t snr_mode, float target_snr_db, uint64_t seed, void* out_host,
                                   int32_t out_dtype, int32_t out_frames, int32_t* status_host) {
  auto bad = [&](const char* m) { set_error(std::string("asr_mfcc_batch_host: ") + m); return ASR_ERR_INVALID; };
  if (!plan) return bad("null plan");
  if (n_clips < 0) return bad("negative n_clips");
  if (n_clips == 0) return ASR_OK;
  if (!audio_host || !offsets_host || !lengths_host || !out_host) return bad("null pointer");
  if (dtype < ASR_I16 || dtype > ASR_F64) return bad("bad dtype");
  if (out_dtype != ASR_F32 && out_dtype != ASR_F64) return bad("bad out_dtype");
  if (out_frames < 1) return bad("out_frames must be >= 1");
  if (snr_mode != 0 && snr_mode != 1) return bad("snr_mode must be 0 or 1");
  if (snr_mode == 1 && dtype == ASR_F64) return bad("SNR mode needs ASR_I16 or ASR_F32 audio");
  const size_t esz = dtype == ASR_I16 ? 2 : (dtype == ASR_F32 ? 4 : 8);
  const size_t osz = out_dtype == ASR_F64 ? 8 : 4;
  const int rows = asr_plan_feature_rows(plan);
  const size_t out_per_clip = static_cast<size_t>(rows) * out_frames;
  // chunks of consecutive clips: <= 32 MiB of samples (ASR_B200_HOST_CHUNK_MIB) and <= 4096 clips; offsets must be ascending & disjoint
  static const int64_t chunk_mib = [] { const char* e = std::getenv("ASR_B200_HOST_CHUNK_MIB"); const int v = e ? std::atoi(e) : 0; return static_cast<int64_t>(v > 0 ? v : 32); }();   // 32 MiB: 5.17 ms per 8192 one-second clips against 5.40 (64) / 5.18 (16) / 5.58 (8), scripts/host_call_time.py
  const int64_t kChunkSamples = (chunk_mib << 20) / static_cast<int64_t>(esz);
  const int kChunkClips = 4096;
  std::vector<int> cuts{0};
  {
    int64_t s = 0;
    int c = 0;
    for (int i = 0; i < n_clips; ++i) {
      if (lengths_host[i] < 0) return bad("negative clip length");
      if (i > 0 && offsets_host[i] < offsets_host[i - 1] + lengths_host[i - 1])
        return bad("offsets must be ascending and clips must not overlap");
      if (c > 0 && (c >= kChunkClips || s + lengths_host[i] > kChunkSamples)) { cuts.push_back(i); s = 0; c = 0; }
      s += lengths_host[i];
      ++c;
    }
    cuts.push_back(n_clips);
  }
  size_t max_span = 0;
  int max_clips = 0, max_len = 0;
  for (size_t k = 0; k + 1 < cuts.size(); ++k) {
    const int a = cuts[k], b = cuts[k + 1];
    const size_t span = static_cast<size_t>(offsets_host[b - 1] + lengths_host[b - 1] - offsets_host[a]);
    max_span = std::max(max_span, span);
    max_clips = std::max(max_clips, b - a);
  }
  for (int i = 0; i < n_clips; ++i) max_len = std::max(max_len, lengths_host[i]);
  static std::mutex create_mu;
  HostState* hs;
  {
    std::lock_guard<std::mutex> g(create_mu);
    if (!plan->host_state) const_cast<asr_plan*>(plan)->host_state = new HostState();
    hs = static_cast<HostState*>(plan->host_state);
  }
  std::lock_guard<std::mutex> lock(hs->mu);
  Slot* slots = hs->slots;
  int rc = ASR_OK;
  auto fail = [&](cudaError_t e, const char* what) { rc = cuda_fail(e, what); };
  const size_t ws_bytes = asr_mfcc_workspace_bytes(plan, max_clips, max_len);
  for (int s = 0; s < 2 && rc == ASR_OK; ++s) {
    Slot& sl = slots[s];
    cudaError_t e = cudaSuccess;
    if (!sl.st) e = cudaStreamCreateWithFlags(&sl.st, cudaStreamNonBlocking);
    if (e == cudaSuccess && !sl.uploaded) e = cudaEventCreateWithFlags(&sl.uploaded, cudaEventDisableTiming);
    if (e == cudaSuccess) e = Slot::grow(&sl.audio, &sl.audio_cap, std::max<size_t>(16, max_span * esz));
    if (e == cudaSuccess && static_cast<size_t>(max_clips) > sl.clip_cap) {
      size_t c0 = 0, c1 = 0, c2 = 0;                       // the three per-clip arrays grow together
      if (sl.offsets) cudaFree(sl.offsets);
      if (sl.lengths) cudaFree(sl.lengths);
      if (sl.status) cudaFree(sl.status);
      sl.offsets = nullptr; sl.lengths = nullptr; sl.status = nullptr; sl.clip_cap = 0;
      e = Slot::grow(&sl.offsets, &c0, sizeof(long long) * max_clips);
      if (e == cudaSuccess) e = Slot::grow(&sl.lengths, &c1, sizeof(int) * max_clips);
      if (e == cudaSuccess) e = Slot::grow(&sl.status, &c2, sizeof(int) * max_clips);
      if (e == cudaSuccess) sl.clip_cap = static_cast<size_t>(max_clips);
    }
    if (e == cudaSuccess) e = Slot::grow(&sl.out, &sl.out_cap, out_per_clip * osz * max_clips);
    if (e == cudaSuccess && snr_mode) {
      if (static_cast<size_t>(max_clips) > sl.snr_cap) {
        size_t c0 = 0, c1 = 0;
        if (sl.power) cudaFree(sl.power);
        if (sl.sigma) cudaFree(sl.sigma);
        sl.power = nullptr; sl.sigma = nullptr; sl.snr_cap = 0;
        e = Slot::grow(&sl.power, &c0, sizeof(float) * max_clips);
        if (e == cudaSuccess) e = Slot::grow(&sl.sigma, &c1, sizeof(double) * max_clips);
        if (e == cudaSuccess) sl.snr_cap = static_cast<size_t>(max_clips);
      }
      if (e == cudaSuccess) e = Slot::grow(&sl.z, &sl.z_cap, sizeof(double) * std::max<size_t>(1, max_span));
    }
    if (e == cudaSuccess && ws_bytes) e = Slot::grow(&sl.ws, &sl.ws_cap, ws_bytes);
    if (e == cudaSuccess && static_cast<size_t>(max_clips) > sl.reb_cap) {
      if (sl.reb_pinned) cudaFreeHost(sl.reb_pinned);
      if (sl.len_pinned) cudaFreeHost(sl.len_pinned);
      sl.reb_pinned = nullptr; sl.len_pinned = nullptr; sl.reb_cap = 0;
      e = cudaHostAlloc(reinterpret_cast<void**>(&sl.reb_pinned), sizeof(long long) * max_clips, cudaHostAllocDefault);
      if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&sl.len_pinned), sizeof(int) * max_clips, cudaHostAllocDefault);
      if (e == cudaSuccess) sl.reb_cap = static_cast<size_t>(max_clips);
    }
    if (e != cudaSuccess) fail(e, "asr_mfcc_batch_host (allocation)");
  }
  if (rc == ASR_OK && status_host && static_cast<size_t>(n_clips) > hs->status_cap) {
    if (hs->status_pinned) cudaFreeHost(hs->status_pinned);
    hs->status_pinned = nullptr; hs->status_cap = 0;
    const cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&hs->status_pinned), sizeof(int) * n_clips, cudaHostAllocDefault);
    if (e == cudaSuccess) hs->status_cap = static_cast<size_t>(n_clips); else fail(e, "asr_mfcc_batch_host (allocation)");
  }
  for (size_t k = 0; k + 1 < cuts.size() && rc == ASR_OK; ++k) {
    Slot& sl = slots[k & 1];
    const int a = cuts[k], b = cuts[k + 1], nc = b - a;
    const int64_t first = offsets_host[a];
    const size_t span = static_cast<size_t>(offsets_host[b - 1] + lengths_host[b - 1] - first);
    // the slot's previous chunk (k-2) must have left the pinned descriptor buffer; its kernels and copies are ordered on the
    // slot's stream, so nothing else has to be waited for here: chunk k's upload overlaps chunk k-1's kernels / download
    cudaError_t e = k >= 2 ? cudaEventSynchronize(sl.uploaded) : cudaSuccess;
    if (e != cudaSuccess) { fail(e, "event sync"); break; }
    for (int i = 0; i < nc; ++i) { sl.reb_pinned[i] = offsets_host[a + i] - first; sl.len_pinned[i] = lengths_host[a + i]; }
    e = cudaMemcpyAsync(sl.audio, static_cast<const char*>(audio_host) + first * esz, span * esz,
                        cudaMemcpyHostToDevice, sl.st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sl.offsets, sl.reb_pinned, sizeof(long long) * nc, cudaMemcpyHostToDevice, sl.st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sl.lengths, sl.len_pinned, sizeof(int) * nc, cudaMemcpyHostToDevice, sl.st);
    if (e == cudaSuccess) e = cudaEventRecord(sl.uploaded, sl.st);
    if (e != cudaSuccess) { fail(e, "H2D copy"); break; }
    asr_noise nz;
    std::memset(&nz, 0, sizeof(nz));
    if (snr_mode) {
      rc = asr_clip_power(sl.audio, dtype, reinterpret_cast<const int64_t*>(sl.offsets), sl.lengths, nc, sl.power, sl.st);
      if (rc == ASR_OK) rc = asr_snr_sigma(sl.power, target_snr_db, sl.sigma, nc, sl.st);
      if (rc == ASR_OK) rc = asr_randn_f64(seed, static_cast<uint64_t>(first), static_cast<int64_t>(span), sl.z, sl.st);
      if (rc != ASR_OK) break;
      nz.mode = ASR_NOISE_WHITE; nz.z_dev = sl.z; nz.sigma_dev = sl.sigma;
    }
    rc = asr_mfcc_batch(plan, sl.audio, dtype, reinterpret_cast<const int64_t*>(sl.offsets), sl.lengths, nc, max_len,
                        snr_mode ? &nz : nullptr, sl.out, out_dtype, out_frames, sl.status, sl.ws, sl.ws ? sl.ws_cap : 0, sl.st);
    if (rc != ASR_OK) break;
    e = cudaMemcpyAsync(static_cast<char*>(out_host) + static_cast<size_t>(a) * out_per_clip * osz, sl.out,
                        out_per_clip * osz * nc, cudaMemcpyDeviceToHost, sl.st);
    if (e == cudaSuccess && status_host)
      e = cudaMemcpyAsync(hs->status_pinned + a, sl.status, sizeof(int) * nc, cudaMemcpyDeviceToHost, sl.st);
    if (e != cudaSuccess) { fail(e, "D2H copy"); break; }
  }
  for (int s = 0; s < 2; ++s)
    if (slots[s].st) {
      const cudaError_t e = cudaStreamSynchronize(slots[s].st);
      if (e != cudaSuccess && rc == ASR_OK) fail(e, "final sync");
    }
  if (rc == ASR_OK && status_host) std::memcpy(status_host, hs->status_pinned, sizeof(int) * n_clips);
  return rc;
}
