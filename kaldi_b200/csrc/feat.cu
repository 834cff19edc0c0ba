// feat.cu — B200-native batched MFCC / fbank extraction + online CMVN (sm_100a).
//
// Replaces, behind the C-ABI, the per-frame CPU chain of the reference
//   ExtractWindow/ProcessWindow   feat/feature-window.cc:137-224
//   SplitRadixRealFft::Compute    matrix/srfft.cc:355-432
//   ComputePowerSpectrum          feat/feature-functions.cc:29-51
//   MelBanks::Compute             feat/mel-computations.cc:226-251
//   MfccComputer::Compute         feat/feature-mfcc.cc:28-80
//   FbankComputer::Compute        feat/feature-fbank.cc:72-123
//   OnlineCmvn::GetFrame          feat/online-feature.cc:421-452 (+ cmvn.cc:64-115)
// and the 10-kernel + cuFFT + cuBLAS sequence of cudafeat:
// here the whole chain framing -> DC removal -> pre-emphasis -> window -> 512-pt
// real FFT -> power -> mel -> log -> DCT -> lifter is ONE kernel, one warp per
// frame, the frame living in shared memory from the first load to the last
// store (800 algorithmic bytes per frame: 640 B of new samples read, 160 B
// written).  Tables (window, twiddles, mel weights, DCT, lifter) are built on
// the host exactly as the reference constructors do and are staged in shared
// memory once per CTA.
//
// The online CMVN is a separate, tiny kernel: one thread per (lane, dim) walks
// the frames in order with double-precision sliding-window statistics in the
// same operation order as OnlineCmvn::ComputeStatsForFrame, so its output is
// bit-identical to the reference when frames are requested in order.

#include <cmath>
#include <vector>

#include "common.cuh"
#include "feat_kernels.cuh"

namespace b2k {

#define FEAT_NFFT 512
#define FEAT_MAX_BINS 64
#define FEAT_MAX_MELW 1024
#define FEAT_WARPS 8
#define FEAT_FRAMES_PER_CTA 64

struct FeatTables {               // device pointers
  const float *window;            // [frame_length]
  const float2 *tw_half;          // [NFFT/4]  exp(-2 pi i k / (NFFT/2))
  const float2 *tw_full;          // [NFFT/2]  exp(-2 pi i k / NFFT)
  const int *mel_first, *mel_len, *mel_off;   // [num_bins]
  const float *mel_w;             // [mel_w_total]
  const float *dct;               // [num_ceps * num_bins]
  const float *lifter;            // [num_ceps]
  const float *equal_loudness;    // plp: [num_bins] (GetEqualLoudnessVector, mel-computations.cc:301-313)
  const float *idft;              // plp: [(lpc_order + 1) x (num_bins + 2)] (InitIdftBases, feature-functions.cc:188-203)
};

struct FeatParams {
  FeatTables t;
  int frame_length, frame_shift, num_bins, num_ceps, mel_w_total;
  int feature_type, remove_dc, snip_edges, use_energy, raw_energy, htk_compat, use_log_fbank,
      use_power, htk_mode, use_lifter;
  float preemph, energy_floor_log, has_energy_floor;
  int dim;
  int lpc_order; float compress_factor, cepstral_scale;     // plp
};

struct FeatLane {                 // per lane descriptor (device array)
  const float *wave;              // utterance samples (whole utterance so far), int16-range floats; null when wave16 is given
  const int16_t *wave16;          // the same as 16-bit PCM (the format audio arrives in: 320 instead of 640 bytes per frame)
  int num_samples;                // valid samples in wave
  int first_frame, num_frames;    // frames to compute in this call
  float *out;                     // &feats[first_frame][0] is out + first_frame*row_stride
  int row_stride;
};

__device__ __forceinline__ int bitrev8(int x) { return (int)(__brev((unsigned)x) >> 24); }

__global__ void __launch_bounds__(FEAT_WARPS * 32)
feat_kernel(FeatParams p, const FeatLane *lanes, int frames_per_cta, int stage_cap_bytes) {
  extern __shared__ float smem[];
  // layout
  float *s_window = smem;                                   // frame_length (<=512)
  float2 *s_tw_half = reinterpret_cast<float2 *>(s_window + FEAT_NFFT);      // 128
  float2 *s_tw_full = s_tw_half + FEAT_NFFT / 4;            // 256
  float *s_melw = reinterpret_cast<float *>(s_tw_full + FEAT_NFFT / 2);      // FEAT_MAX_MELW
  int *s_mel_first = reinterpret_cast<int *>(s_melw + FEAT_MAX_MELW);        // 64
  int *s_mel_len = s_mel_first + FEAT_MAX_BINS;
  int *s_mel_off = s_mel_len + FEAT_MAX_BINS;
  float *s_dct = reinterpret_cast<float *>(s_mel_off + FEAT_MAX_BINS);       // 64*64
  float *s_lifter = s_dct + FEAT_MAX_BINS * FEAT_MAX_BINS;  // 64
  float *s_frames = s_lifter + FEAT_MAX_BINS;               // FEAT_WARPS * (NFFT + 2)
  float *s_mel = s_frames + FEAT_WARPS * (FEAT_NFFT + 2);   // FEAT_WARPS * 64

  const int tid = threadIdx.x, lane_id = tid & 31, warp = tid >> 5;
  for (int i = tid; i < p.frame_length; i += blockDim.x) s_window[i] = p.t.window[i];
  for (int i = tid; i < FEAT_NFFT / 4; i += blockDim.x) s_tw_half[i] = p.t.tw_half[i];
  for (int i = tid; i < FEAT_NFFT / 2; i += blockDim.x) s_tw_full[i] = p.t.tw_full[i];
  for (int i = tid; i < p.mel_w_total; i += blockDim.x) s_melw[i] = p.t.mel_w[i];
  for (int i = tid; i < p.num_bins; i += blockDim.x) {
    s_mel_first[i] = p.t.mel_first[i]; s_mel_len[i] = p.t.mel_len[i]; s_mel_off[i] = p.t.mel_off[i];
  }
  for (int i = tid; i < p.num_ceps * p.num_bins; i += blockDim.x) s_dct[i] = p.t.dct[i];
  for (int i = tid; i < p.num_ceps; i += blockDim.x) s_lifter[i] = p.use_lifter ? p.t.lifter[i] : 1.0f;
  __syncthreads();

  const FeatLane L = lanes[blockIdx.y];
  float *buf = s_frames + warp * (FEAT_NFFT + 2);      // real samples, then complex z in place
  float2 *z = reinterpret_cast<float2 *>(buf);
  float *mel = s_mel + warp * FEAT_MAX_BINS;
  const int f_begin = blockIdx.x * frames_per_cta;
  const int f_end = min(f_begin + frames_per_cta, L.num_frames);
  const int NL = p.frame_length;
  const float FLT_EPS = 1.1920928955078125e-07f;
  if (f_begin >= f_end) return;                         // (uniform: nothing below is reached by a part of the CTA only)

  // ---- the samples of this CTA's frames: ONE bulk copy (TMA, cp.async.bulk) of the contiguous span they cover into shared
  //      memory, completion on an mbarrier; the frames overlap by 60 %, so every sample is fetched from HBM / L2 once per CTA
  //      instead of 2.5 times.  The span is cut to 16-byte boundaries of the source; what falls outside it (the ragged ends,
  //      reflected samples of snip_edges = false) is read directly.
  __shared__ __align__(8) unsigned long long s_bar;
  unsigned char *s_stage = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<unsigned long long>(s_mel + FEAT_WARPS * FEAT_MAX_BINS) + 15ull) & ~15ull);      // 16 spare bytes are reserved
  long long st_lo = 0, st_hi = 0;                        // staged samples [st_lo, st_hi)
  {
    const int esz = L.wave16 ? 2 : 4;
    const unsigned char *base = L.wave16 ? reinterpret_cast<const unsigned char *>(L.wave16) : reinterpret_cast<const unsigned char *>(L.wave);
    const long long fr0 = L.first_frame + f_begin, fr1 = L.first_frame + f_end - 1;
    long long lo = p.snip_edges ? fr0 * p.frame_shift : (long long)p.frame_shift * fr0 + p.frame_shift / 2 - NL / 2;
    long long hi = (p.snip_edges ? fr1 * p.frame_shift : (long long)p.frame_shift * fr1 + p.frame_shift / 2 - NL / 2) + NL;
    lo = max(lo, 0LL); hi = min(hi, (long long)L.num_samples);
    unsigned long long a0 = 0, a1 = 0;
    if (hi > lo) {
      a0 = (reinterpret_cast<unsigned long long>(base) + (unsigned long long)lo * esz + 15ull) & ~15ull;     // first 16-byte boundary inside
      a1 = (reinterpret_cast<unsigned long long>(base) + (unsigned long long)hi * esz) & ~15ull;             // last one inside
    }
    unsigned bytes = 0;
    if (a1 > a0 && a1 - a0 <= (unsigned long long)stage_cap_bytes) bytes = (unsigned)(a1 - a0);
    if (bytes) {
      st_lo = (long long)((a0 - reinterpret_cast<unsigned long long>(base)) / esz);
      st_hi = st_lo + bytes / esz;
    }
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (bytes) {
      if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"((unsigned)__cvta_generic_to_shared(s_stage)), "l"(a0), "r"(bytes), "r"(bar) : "memory");
      }
      unsigned ok = 0;
      do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar) : "memory");
      } while (!ok);
    }
  }
  const float *stage_f = reinterpret_cast<const float *>(s_stage);
  const int16_t *stage_h = reinterpret_cast<const int16_t *>(s_stage);

  for (int fr = f_begin + warp; fr < f_end; fr += FEAT_WARPS) {
    const int frame = L.first_frame + fr;
    // FirstSampleOfFrame (feature-window.cc:30-40)
    long long start = p.snip_edges ? (long long)frame * p.frame_shift
                                   : (long long)p.frame_shift * frame + p.frame_shift / 2 - NL / 2;
    // ---- gather with reflection (:202-214), DC removal (:146-147)
    float v[16];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 16; k++) {
      int i = lane_id + 32 * k;
      float x = 0.f;
      if (i < NL) {
        long long s = start + i;
        if (s < 0 || s >= L.num_samples) {
          int n = L.num_samples;
          while (s < 0 || s >= n) { if (s < 0) s = -s - 1; else s = 2LL * n - 1 - s; }
        }
        if (s >= st_lo && s < st_hi) x = L.wave16 ? (float)stage_h[s - st_lo] : stage_f[s - st_lo];
        else x = L.wave16 ? (float)L.wave16[s] : L.wave[s];
        sum += x;
      }
      v[k] = x;
    }
    if (p.remove_dc) {
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      float off = -sum / (float)NL;
#pragma unroll
      for (int k = 0; k < 16; k++) if (lane_id + 32 * k < NL) v[k] += off;
    }
    float log_energy = 0.f;
    if (p.use_energy && p.raw_energy) {                     // :149-153
      float e = 0.f;
#pragma unroll
      for (int k = 0; k < 16; k++) e += v[k] * v[k];
      for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
      log_energy = logf(fmaxf(e, FLT_EPS));
    }
    // ---- pre-emphasis (:100-107) needs x[i-1]: stage in smem
#pragma unroll
    for (int k = 0; k < 16; k++) buf[lane_id + 32 * k] = v[k];
    __syncwarp();
    if (p.preemph != 0.f) {
#pragma unroll
      for (int k = 0; k < 16; k++) {
        int i = lane_id + 32 * k;
        if (i < NL) {
          float prev = (i == 0) ? v[k] : buf[i - 1];
          v[k] = v[k] - p.preemph * prev;
        }
      }
    }
    // ---- window (:159)
    float e_win = 0.f;
#pragma unroll
    for (int k = 0; k < 16; k++) {
      int i = lane_id + 32 * k;
      if (i < NL) v[k] *= s_window[i]; else v[k] = 0.f;
      e_win += v[k] * v[k];
    }
    if (p.use_energy && !p.raw_energy) {                    // feature-mfcc.cc:37-39
      for (int o = 16; o > 0; o >>= 1) e_win += __shfl_xor_sync(0xffffffffu, e_win, o);
      log_energy = logf(fmaxf(e_win, FLT_EPS));
    }
    __syncwarp();
    // ---- real FFT of size 512 via complex FFT of size 256 on z[n] = x[2n] + i x[2n+1],
    //      radix-2 DIT, input scattered to bit-reversed positions
#pragma unroll
    for (int k = 0; k < 16; k++) {
      int i = lane_id + 32 * k;                 // sample index; complex index n = i>>1, part = i&1
      int n = i >> 1;
      buf[2 * bitrev8(n) + (i & 1)] = v[k];
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 8; s++) {
      const int half = 1 << s;
      const int tw_stride = (FEAT_NFFT / 4) >> s;            // 128 / half
#pragma unroll
      for (int q = 0; q < 4; q++) {
        int j = lane_id + 32 * q;                           // butterfly 0..127
        int pos = j & (half - 1);
        int i0 = ((j >> s) << (s + 1)) + pos;
        int i1 = i0 + half;
        float2 w = s_tw_half[pos * tw_stride];
        float2 a = z[i0], b = z[i1];
        float2 t = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
        z[i0] = make_float2(a.x + t.x, a.y + t.y);
        z[i1] = make_float2(a.x - t.x, a.y - t.y);
      }
      __syncwarp();
    }
    // ---- post-process to the real spectrum and take the power (feature-functions.cc:29-51)
    //      X[k] = (A+B)/2 - (i/2) W^k (A-B),  A = Z[k], B = conj(Z[256-k])
    float pw[9];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      int k = lane_id + 32 * q;                             // 0..255
      float pk;
      if (k == 0) {
        float2 z0 = z[0];
        float x0 = z0.x + z0.y;
        pk = x0 * x0;
      } else {
        float2 A = z[k], Zc = z[256 - k];
        float2 B = make_float2(Zc.x, -Zc.y);
        float2 sm = make_float2(0.5f * (A.x + B.x), 0.5f * (A.y + B.y));
        float2 df = make_float2(0.5f * (A.x - B.x), 0.5f * (A.y - B.y));
        float2 w = s_tw_full[k];
        float2 wd = make_float2(w.x * df.x - w.y * df.y, w.x * df.y + w.y * df.x);
        // -i * wd = (wd.y, -wd.x)
        float re = sm.x + wd.y, im = sm.y - wd.x;
        pk = re * re + im * im;
      }
      pw[q] = pk;
    }
    {
      float2 z0 = z[0];
      float xn = z0.x - z0.y;
      pw[8] = xn * xn;                                      // bin 256
    }
    __syncwarp();
    float *pspec = buf;                                     // reuse: 257 floats
#pragma unroll
    for (int q = 0; q < 8; q++) {
      float val = pw[q];
      if (p.feature_type == 1 && !p.use_power) val = sqrtf(val);
      pspec[lane_id + 32 * q] = val;
    }
    if (lane_id == 0) pspec[256] = (p.feature_type == 1 && !p.use_power) ? sqrtf(pw[8]) : pw[8];
    __syncwarp();
    // ---- mel filterbank (mel-computations.cc:226-251), floor, log
    for (int b = lane_id; b < p.num_bins; b += 32) {
      const float *w = s_melw + s_mel_off[b];
      const float *ps = pspec + s_mel_first[b];
      int len = s_mel_len[b];
      float e = 0.f;
      for (int i = 0; i < len; i++) e = fmaf(w[i], ps[i], e);
      if (p.htk_mode && e < 1.0f) e = 1.0f;
      if (p.feature_type == 0 || (p.feature_type == 1 && p.use_log_fbank)) e = logf(fmaxf(e, FLT_EPS));
      if (p.feature_type == 2) e = powf(e * __ldg(&p.t.equal_loudness[b]), p.compress_factor);   // feature-plp.cc:140-142
      mel[b] = e;
    }
    __syncwarp();
    float *orow = L.out + (size_t)frame * L.row_stride;
    if (p.use_energy && p.has_energy_floor != 0.f && log_energy < p.energy_floor_log)
      log_energy = p.energy_floor_log;
    if (p.feature_type == 2) {
      // PlpComputer::Compute (feature-plp.cc:144-182): autocorrelation of the compressed spectrum with its first and last
      // band repeated, Durbin's recursion, LPC -> cepstrum, lifter, scale, C0 / energy, HTK order.  The frame's scratch
      // (buf is free again) holds the 13 + 12 + 12 + 12 values; the recursions are short and sequential: lane 0 runs them.
      const int N2 = p.num_bins + 2, LO = p.lpc_order;
      float *ac = buf, *lpc = buf + 64, *tmp = buf + 128, *cep = buf + 192;
      for (int k = lane_id; k <= LO; k += 32) {
        const float *row = p.t.idft + (size_t)k * N2;
        float a = __ldg(&row[0]) * mel[0];                                  // AddMatVec: a dot product per row, in order
        for (int j = 1; j <= p.num_bins; j++) a = fmaf(__ldg(&row[j]), mel[j - 1], a);
        a = fmaf(__ldg(&row[N2 - 1]), mel[p.num_bins - 1], a);
        ac[k] = a;
      }
      __syncwarp();
      if (lane_id == 0) {
        float E = ac[0];                                                    // Durbin (mel-computations.cc:266-297)
        for (int i = 0; i < LO; i++) {
          float ki = ac[i + 1];
          for (int j = 0; j < i; j++) ki += lpc[j] * ac[i - j];
          ki = ki / E;
          float c = 1.0f - ki * ki;
          if (c < 1.0e-5f) c = 1.0e-5f;
          E *= c;
          tmp[i] = -ki;
          for (int j = 0; j < i; j++) tmp[j] = lpc[j] - ki * lpc[i - j - 1];
          for (int j = 0; j <= i; j++) lpc[j] = tmp[j];
        }
        float res = -logf(1.0f / E);                                        // ComputeLpc (:317-329)
        res = fmaxf(res, 1.17549435e-38f);                                  // std::numeric_limits<float>::min() (:161-162)
        for (int i = 0; i < LO; i++) {                                      // Lpc2Cepstrum (:300-309), the sum in double
          double sum = 0.0;
          for (int j = 0; j < i; j++) sum += (double)((float)(i - j) * lpc[j] * cep[i - j - 1]);
          cep[i] = (float)((double)(-lpc[i]) - sum / (double)(float)(i + 1));
        }
        for (int c = 0; c < p.num_ceps; c++) {
          float v = (c == 0) ? res : cep[c - 1];
          v *= s_lifter[c];
          if (p.cepstral_scale != 1.0f) v *= p.cepstral_scale;
          if (p.use_energy && c == 0) v = log_energy;
          if (p.htk_compat) { if (c == 0) orow[p.num_ceps - 1] = v; else orow[c - 1] = v; }
          else orow[c] = v;
        }
      }
    } else if (p.feature_type == 1) {
      // fbank layout (feature-fbank.cc:103-122)
      int mel_offset = (p.use_energy && !p.htk_compat) ? 1 : 0;
      for (int b = lane_id; b < p.num_bins; b += 32) orow[mel_offset + b] = mel[b];
      if (p.use_energy && lane_id == 0) orow[p.htk_compat ? p.num_bins : 0] = log_energy;
    } else {
      // DCT + lifter (feature-mfcc.cc:58-63), energy (:65-69), htk_compat reorder (:71-80)
      for (int c = lane_id; c < p.num_ceps; c += 32) {
        const float *drow = s_dct + c * p.num_bins;
        float acc = 0.f;
        for (int b = 0; b < p.num_bins; b++) acc = fmaf(drow[b], mel[b], acc);
        acc *= s_lifter[c];
        if (p.use_energy && c == 0) acc = log_energy;
        if (p.htk_compat) {
          if (c == 0) { if (!p.use_energy) acc *= 1.41421356237309504880f; orow[p.num_ceps - 1] = acc; }
          else orow[c - 1] = acc;
        } else {
          orow[c] = acc;
        }
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ online CMVN (structs in feat_kernels.cuh)
// one thread per (lane, dim); count column handled redundantly by every thread
__global__ void cmvn_kernel(CmvnParams p, const CmvnLane *lanes) {
  const CmvnLane L = lanes[blockIdx.x];
  const int d = threadIdx.x;
  const int D = p.dim;
  const bool act = d < D;
  const double *spk = L.speaker ? L.speaker : p.speaker_stats;
  double s0 = 0.0, s1 = 0.0, cnt = 0.0, sp0 = 0.0, sp1 = 0.0, spc = 0.0, sp1c = 0.0;
  if (act) {
    s0 = L.state[d]; s1 = L.state[(D + 1) + d]; cnt = L.state[D];
    if (spk) { sp0 = spk[d]; sp1 = spk[(D + 1) + d]; spc = spk[D]; sp1c = spk[2 * D + 1]; }
  }
  __syncthreads();                 // speaker_out may be the array `speaker` was read from
  if (!act) return;
  const double g0 = p.global_stats ? p.global_stats[d] : 0.0, g1 = p.global_stats ? p.global_stats[(D + 1) + d] : 0.0,
               gc = p.global_stats ? p.global_stats[D] : 0.0;
  // OnlineCmvn::GetState (online-feature.cc:278-300): the speaker stats plus every frame of this utterance, in double
  double a0 = sp0, a1 = sp1, ac = spc;
  for (int i = 0; i < L.num_frames; i++) {
    const int t = L.first_frame + i;
    const float xf = L.in[(size_t)t * L.in_stride + d];
    const double x = (double)xf;
    // ComputeStatsForFrame (online-feature.cc:346-366)
    s0 += x;
    if (p.normalize_variance) s1 = __dadd_rn(s1, __dmul_rn(x, x));   // AddVec2: no FMA contraction on the CPU
    cnt += 1.0;
    const int prev = t - p.cmn_window;
    if (prev >= 0) {
      const double y = (double)L.in[(size_t)prev * L.in_stride + d];
      s0 -= y;
      if (p.normalize_variance) s1 = __dadd_rn(s1, -__dmul_rn(y, y));
      cnt -= 1.0;
    }
    // SmoothOnlineCmvnStats (:372-419)
    double m0 = s0, m1 = s1, c = cnt;
    if (L.speaker_out) { a0 += x; a1 = __dadd_rn(a1, __dmul_rn(x, x)); ac += 1.0; }
    if (c < (double)p.cmn_window) {
      if (spk) {
        double cfs = (double)p.cmn_window - c;
        if (cfs > (double)p.speaker_frames) cfs = (double)p.speaker_frames;
        if (cfs > spc) cfs = spc;
        if (cfs > 0.0) { double a = cfs / spc; m0 = __dadd_rn(m0, __dmul_rn(a, sp0)); m1 = __dadd_rn(m1, __dmul_rn(a, sp1)); c = __dadd_rn(c, __dmul_rn(a, spc)); }
      }
      if (c < (double)p.cmn_window && p.global_stats) {
        double cfg = (double)p.cmn_window - c;
        if (cfg > (double)p.global_frames) cfg = (double)p.global_frames;
        if (cfg > 0.0) { double a = cfg / gc; m0 = __dadd_rn(m0, __dmul_rn(a, g0)); m1 = __dadd_rn(m1, __dmul_rn(a, g1)); c = __dadd_rn(c, __dmul_rn(a, gc)); }
      }
    }
    // ApplyCmvn (transform/cmvn.cc:64-115)
    float y = xf;
    if (p.normalize_mean) {
      if (!p.normalize_variance) {
        const float alpha = (float)(-1.0 / c);              // AddVec alpha is BaseFloat
        const float off = (float)((double)alpha * m0);
        y = xf + off;
      } else {
        double mean = m0 / c;
        double var = __dadd_rn(m1 / c, -__dmul_rn(mean, mean));
        if (var < 1.0e-20) var = 1.0e-20;
        double scale = 1.0 / sqrt(var);
        double offset = -(mean * scale);
        y = __fadd_rn(__fmul_rn(xf, (float)scale), (float)offset);   // MulColsVec then AddVecToRows
      }
    }
    L.out[(size_t)t * L.out_stride + d] = y;
  }
  L.state[d] = s0; L.state[(D + 1) + d] = s1;
  if (d == 0) L.state[D] = cnt;
  if (L.speaker_out) {
    // OnlineIvectorExtractorAdaptationState::LimitFrames, CMVN half (online-ivector-feature.cc:113-118): BaseFloat arithmetic
    double sc = 1.0;
    const float count = (float)ac;
    if (L.max_remembered_frames >= 0.0f && count > L.max_remembered_frames) sc = (double)(L.max_remembered_frames / count);
    L.speaker_out[d] = a0 * sc; L.speaker_out[(D + 1) + d] = a1 * sc;
    if (d == 0) { L.speaker_out[D] = ac * sc; L.speaker_out[2 * D + 1] = sp1c * sc; }
  }
}

int launch_cmvn(const CmvnParams &cp, const CmvnLane *d_lanes, int num_lanes, cudaStream_t st) {
  int threads = ((cp.dim + 31) / 32) * 32;
  cmvn_kernel<<<num_lanes, threads, 0, st>>>(cp, d_lanes);
  B2K_LAUNCH_CHECK();
  return B2K_OK;
}

}  // namespace b2k

using namespace b2k;

struct b2k_feat {
  b2k_feat_cfg cfg;
  FeatParams p;
  std::vector<void *> allocs;
  FeatLane *d_lanes = nullptr, *h_lanes = nullptr;
  CmvnLane *d_clanes = nullptr, *h_clanes = nullptr;
  int max_lanes = 0;
  cudaEvent_t staging_free = nullptr;
  size_t smem_bytes = 0, stage_bytes = 0;
  int frame_length = 0, frame_shift = 0;
};

static inline float melscale_f(float f) { return 1127.0f * logf(1.0f + f / 700.0f); }   // mel-computations.h:85

extern "C" {

void b2k_feat_cfg_default(b2k_feat_cfg *c) {
  if (!c) return;
  memset(c, 0, sizeof(*c));
  c->feature_type = 0; c->samp_freq = 16000.f; c->frame_shift_ms = 10.f; c->frame_length_ms = 25.f;
  c->dither = 0.f; c->preemph_coeff = 0.97f; c->remove_dc_offset = 1; c->round_to_power_of_two = 1;
  c->snip_edges = 1; c->window_type = 0; c->num_bins = 40; c->low_freq = 20.f; c->high_freq = -400.f;
  c->num_ceps = 40; c->use_energy = 0; c->energy_floor = 0.f; c->raw_energy = 1; c->cepstral_lifter = 22.f;
  c->htk_compat = 0; c->use_log_fbank = 1; c->use_power = 1; c->htk_mode = 0; c->max_lanes = 1024;
  c->lpc_order = 12; c->compress_factor = 0.33333f; c->cepstral_scale = 1.0f;
}

int b2k_feat_create(const b2k_feat_cfg *cfg, b2k_feat **out) {
  if (!cfg || !out) return set_error(B2K_ERR_INVALID, "b2k_feat_create: bad args");
  if (cfg->dither != 0.0f)
    return set_error(B2K_ERR_INVALID, "dither must be 0: the reference's Dither() is not reproducible (feature-window.cc:90-98)");
  int rc = require_device();
  if (rc) return rc;
  const int NL = (int)(cfg->samp_freq * 0.001f * cfg->frame_length_ms);   // WindowSize feature-window.h:109
  const int shift = (int)(cfg->samp_freq * 0.001f * cfg->frame_shift_ms);
  int padded = 1; while (padded < NL) padded *= 2;
  if (!cfg->round_to_power_of_two || padded != FEAT_NFFT)
    return set_error(B2K_ERR_INVALID, "only a padded window of 512 samples is supported (16 kHz / 25 ms)");
  if (cfg->feature_type < 0 || cfg->feature_type > 2) return set_error(B2K_ERR_INVALID, "feature_type must be 0 (mfcc), 1 (fbank) or 2 (plp)");
  if (cfg->num_bins < 3 || cfg->num_bins > FEAT_MAX_BINS || cfg->num_ceps > cfg->num_bins || cfg->num_ceps < 1)
    return set_error(B2K_ERR_INVALID, "need 3 <= num_bins <= 64 and num_ceps <= num_bins");
  if (cfg->feature_type == 2 && (cfg->lpc_order < 1 || cfg->lpc_order > 63 || cfg->num_ceps > cfg->lpc_order + 1 || cfg->num_ceps < 2))
    return set_error(B2K_ERR_INVALID, "plp: need 1 <= lpc_order <= 63 and 2 <= num_ceps <= lpc_order + 1");   // feature-plp.cc:124
  b2k_feat *f = new b2k_feat();
  f->cfg = *cfg; f->frame_length = NL; f->frame_shift = shift;
  // --- window (FeatureWindowFunction, feature-window.cc:109-135): double math, float storage
  std::vector<float> window(NL);
  {
    double a = 2.0 * M_PI / (NL - 1);
    for (int i = 0; i < NL; i++) {
      double x = (double)i, w;
      switch (cfg->window_type) {
        case 1: w = 0.54 - 0.46 * cos(a * x); break;            // hamming
        case 2: w = 0.5 - 0.5 * cos(a * x); break;              // hanning
        case 3: w = 1.0; break;                                 // rectangular
        default: w = pow(0.5 - 0.5 * cos(a * x), 0.85); break;  // povey
      }
      window[i] = (float)w;
    }
  }
  // --- twiddles
  std::vector<float2> twh(FEAT_NFFT / 4), twf(FEAT_NFFT / 2);
  for (int k = 0; k < FEAT_NFFT / 4; k++) { double a = -2.0 * M_PI * k / (FEAT_NFFT / 2); twh[k] = make_float2((float)cos(a), (float)sin(a)); }
  for (int k = 0; k < FEAT_NFFT / 2; k++) { double a = -2.0 * M_PI * k / FEAT_NFFT; twf[k] = make_float2((float)cos(a), (float)sin(a)); }
  // --- mel banks (MelBanks::MelBanks, mel-computations.cc:33-142, vtln_warp == 1), float arithmetic
  std::vector<int> mfirst(cfg->num_bins), mlen(cfg->num_bins), moff(cfg->num_bins);
  std::vector<float> melw;
  {
    const int num_fft_bins = FEAT_NFFT / 2;
    const float nyquist = 0.5f * cfg->samp_freq;
    const float low = cfg->low_freq;
    const float high = cfg->high_freq > 0.0f ? cfg->high_freq : nyquist + cfg->high_freq;
    if (low < 0.0f || low >= nyquist || high <= 0.0f || high > nyquist || high <= low) {
      delete f; return set_error(B2K_ERR_INVALID, "bad low-freq/high-freq");
    }
    const float fft_bin_width = cfg->samp_freq / FEAT_NFFT;
    const float mel_low = melscale_f(low), mel_high = melscale_f(high);
    const float delta = (mel_high - mel_low) / (cfg->num_bins + 1);
    for (int b = 0; b < cfg->num_bins; b++) {
      float left = mel_low + b * delta, center = mel_low + (b + 1) * delta, right = mel_low + (b + 2) * delta;
      int first = -1, last = -1;
      std::vector<float> w(num_fft_bins, 0.f);
      for (int i = 0; i < num_fft_bins; i++) {
        float mel = melscale_f(fft_bin_width * i);
        if (mel > left && mel < right) {
          w[i] = (mel <= center) ? (mel - left) / (center - left) : (right - mel) / (right - center);
          if (first == -1) first = i;
          last = i;
        }
      }
      if (first == -1) { delete f; return set_error(B2K_ERR_INVALID, "num-mel-bins too large"); }
      if (cfg->htk_mode && b == 0 && mel_low != 0.0f) w[first] = 0.0f;
      mfirst[b] = first; mlen[b] = last + 1 - first; moff[b] = (int)melw.size();
      for (int i = first; i <= last; i++) melw.push_back(w[i]);
    }
    if (melw.size() > FEAT_MAX_MELW) { delete f; return set_error(B2K_ERR_INVALID, "mel weight table too large"); }
  }
  // --- plp: equal-loudness weights at the bands' centre frequencies (GetEqualLoudnessVector, mel-computations.cc:301-313; the
  //     centres as MelBanks::MelBanks leaves them, :99-102) and the inverse-DFT bases (InitIdftBases, feature-functions.cc:188-203)
  std::vector<float> eql(cfg->num_bins, 1.0f), idft(1, 0.0f);
  if (cfg->feature_type == 2) {
    const float nyquist = 0.5f * cfg->samp_freq;
    const float high = cfg->high_freq > 0.0f ? cfg->high_freq : nyquist + cfg->high_freq;
    const float mel_low = melscale_f(cfg->low_freq), mel_high = melscale_f(high);
    const float delta = (mel_high - mel_low) / (cfg->num_bins + 1);
    for (int b = 0; b < cfg->num_bins; b++) {
      const float center_mel = mel_low + (b + 1) * delta;
      const float f0 = 700.0f * (expf(center_mel / 1127.0f) - 1.0f);          // InverseMelScale (mel-computations.h:81)
      const float fsq = f0 * f0;
      const float fsub = (float)(fsq / (fsq + 1.6e5));
      eql[b] = (float)(fsub * fsub * ((fsq + 1.44e6) / (fsq + 9.61e6)));
    }
    const int nb = cfg->lpc_order + 1, dm = cfg->num_bins + 2;
    idft.assign((size_t)nb * dm, 0.0f);
    const float angle = (float)(M_PI / (float)(dm - 1));
    const float scale = (float)(1.0f / (2.0 * (float)(dm - 1)));
    for (int i = 0; i < nb; i++) {
      idft[(size_t)i * dm] = (float)(1.0 * scale);
      const float i_fl = (float)i;
      for (int j = 1; j < dm - 1; j++) idft[(size_t)i * dm + j] = (float)(2.0 * scale * cos(angle * i_fl * (float)j));
      idft[(size_t)i * dm + dm - 1] = (float)(scale * cos(angle * i_fl * (float)(dm - 1)));
    }
  }
  // --- DCT rows (ComputeDctMatrix, matrix-functions.cc:592-608, Real = float) and lifter (mel-computations.cc:253-259)
  std::vector<float> dct((size_t)cfg->num_ceps * cfg->num_bins), lifter(cfg->num_ceps, 1.0f);
  {
    const int N = cfg->num_bins;
    float n0 = std::sqrt(1.0f / (float)N);
    float n1 = std::sqrt(2.0f / (float)N);
    for (int k = 0; k < cfg->num_ceps; k++)
      for (int n = 0; n < N; n++)
        dct[(size_t)k * N + n] = (k == 0) ? n0 : (float)(n1 * std::cos((double)M_PI / N * (n + 0.5) * k));
    if (cfg->cepstral_lifter != 0.0f)
      for (int i = 0; i < cfg->num_ceps; i++)
        lifter[i] = (float)(1.0 + 0.5 * cfg->cepstral_lifter * sin(M_PI * i / cfg->cepstral_lifter));
  }
  auto up = [&](const void *h, size_t bytes, const void **d) -> int {
    void *ptr = nullptr;
    B2K_CUDA_CHECK(cudaMalloc(&ptr, bytes ? bytes : 16));
    f->allocs.push_back(ptr);
    if (bytes) B2K_CUDA_CHECK(cudaMemcpy(ptr, h, bytes, cudaMemcpyHostToDevice));
    *d = ptr;
    return 0;
  };
  FeatParams &p = f->p;
  memset(&p, 0, sizeof(p));
#define UP(field, vec) if ((rc = up((vec).data(), sizeof((vec)[0]) * (vec).size(), (const void **)&p.t.field))) return rc;
  UP(window, window) UP(tw_half, twh) UP(tw_full, twf) UP(mel_first, mfirst) UP(mel_len, mlen) UP(mel_off, moff)
  UP(mel_w, melw) UP(dct, dct) UP(lifter, lifter) UP(equal_loudness, eql) UP(idft, idft)
#undef UP
  p.frame_length = NL; p.frame_shift = shift; p.num_bins = cfg->num_bins; p.num_ceps = cfg->num_ceps;
  p.mel_w_total = (int)melw.size(); p.feature_type = cfg->feature_type; p.remove_dc = cfg->remove_dc_offset;
  p.snip_edges = cfg->snip_edges; p.use_energy = cfg->use_energy; p.raw_energy = cfg->raw_energy;
  p.htk_compat = cfg->htk_compat; p.use_log_fbank = cfg->use_log_fbank; p.use_power = cfg->use_power;
  p.htk_mode = cfg->htk_mode; p.use_lifter = cfg->cepstral_lifter != 0.0f; p.preemph = cfg->preemph_coeff;
  p.has_energy_floor = cfg->energy_floor > 0.0f ? 1.f : 0.f;
  p.energy_floor_log = cfg->energy_floor > 0.0f ? logf(cfg->energy_floor) : 0.f;
  p.dim = cfg->feature_type != 1 ? cfg->num_ceps : cfg->num_bins + (cfg->use_energy ? 1 : 0);
  p.lpc_order = cfg->lpc_order; p.compress_factor = cfg->compress_factor; p.cepstral_scale = cfg->cepstral_scale;
  f->max_lanes = cfg->max_lanes > 0 ? cfg->max_lanes : 1024;
  B2K_CUDA_CHECK(cudaMalloc((void **)&f->d_lanes, sizeof(FeatLane) * f->max_lanes)); f->allocs.push_back(f->d_lanes);
  B2K_CUDA_CHECK(cudaMalloc((void **)&f->d_clanes, sizeof(CmvnLane) * f->max_lanes)); f->allocs.push_back(f->d_clanes);
  B2K_CUDA_CHECK(cudaMallocHost((void **)&f->h_lanes, sizeof(FeatLane) * f->max_lanes));
  B2K_CUDA_CHECK(cudaMallocHost((void **)&f->h_clanes, sizeof(CmvnLane) * f->max_lanes));
  B2K_CUDA_CHECK(cudaEventCreateWithFlags(&f->staging_free, cudaEventDisableTiming));
  f->smem_bytes = sizeof(float) * (FEAT_NFFT + 2 * (FEAT_NFFT / 4) + 2 * (FEAT_NFFT / 2) + FEAT_MAX_MELW +
                                   3 * FEAT_MAX_BINS + FEAT_MAX_BINS * FEAT_MAX_BINS + FEAT_MAX_BINS +
                                   FEAT_WARPS * (FEAT_NFFT + 2) + FEAT_WARPS * FEAT_MAX_BINS);
  // + the staged samples of one CTA: FEAT_FRAMES_PER_CTA frames of fp32 (int16 needs half)
  f->stage_bytes = ((size_t)(FEAT_FRAMES_PER_CTA - 1) * shift + NL) * sizeof(float) + 16;
  f->stage_bytes = (f->stage_bytes + 15) / 16 * 16;
  f->smem_bytes += f->stage_bytes;
  B2K_CUDA_CHECK(cudaFuncSetAttribute(feat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem_bytes));
  *out = f;
  return B2K_OK;
}

int b2k_feat_destroy(b2k_feat *f) {
  if (!f) return B2K_OK;
  cudaDeviceSynchronize();
  for (void *p : f->allocs) cudaFree(p);
  cudaFreeHost(f->h_lanes); cudaFreeHost(f->h_clanes);
  if (f->staging_free) cudaEventDestroy(f->staging_free);
  delete f;
  return B2K_OK;
}

int32_t b2k_feat_dim(const b2k_feat *f) { return f ? f->p.dim : -1; }
float b2k_feat_samp_freq(const b2k_feat *f) { return f ? f->cfg.samp_freq : -1.0f; }

// NumFrames (feat/feature-window.cc:42-87)
int32_t b2k_feat_num_frames(const b2k_feat *f, int64_t num_samples, int32_t flush) {
  if (!f) return -1;
  const int64_t shift = f->frame_shift, length = f->frame_length;
  if (f->cfg.snip_edges) return num_samples < length ? 0 : (int32_t)(1 + (num_samples - length) / shift);
  int32_t n = (int32_t)((num_samples + shift / 2) / shift);
  if (flush) return n;
  int64_t end = (shift * (n - 1) + shift / 2 - length / 2) + length;
  while (n > 0 && end > num_samples) { n--; end -= shift; }
  return n;
}

static int feat_compute_impl(b2k_feat *f, int32_t num_lanes, const float *const *d_wave, const int16_t *const *d_wave16,
                             const int32_t *num_samples, const int32_t *first_frame, const int32_t *num_frames, float *const *d_out,
                             int32_t row_stride, void *stream);

int b2k_feat_compute_batched(b2k_feat *f, int32_t num_lanes, const float *const *d_wave,
                             const int32_t *num_samples, const int32_t *first_frame,
                             const int32_t *num_frames, float *const *d_out, int32_t row_stride,
                             void *stream) {
  return feat_compute_impl(f, num_lanes, d_wave, nullptr, num_samples, first_frame, num_frames, d_out, row_stride, stream);
}

int b2k_feat_compute_batched_i16(b2k_feat *f, int32_t num_lanes, const int16_t *const *d_wave16,
                                 const int32_t *num_samples, const int32_t *first_frame,
                                 const int32_t *num_frames, float *const *d_out, int32_t row_stride,
                                 void *stream) {
  return feat_compute_impl(f, num_lanes, nullptr, d_wave16, num_samples, first_frame, num_frames, d_out, row_stride, stream);
}

static int feat_compute_impl(b2k_feat *f, int32_t num_lanes, const float *const *d_wave, const int16_t *const *d_wave16,
                             const int32_t *num_samples, const int32_t *first_frame, const int32_t *num_frames, float *const *d_out,
                             int32_t row_stride, void *stream) {
  if (!f || num_lanes <= 0 || num_lanes > f->max_lanes || (!d_wave && !d_wave16) || !num_samples || !num_frames || !d_out)
    return set_error(B2K_ERR_INVALID, "b2k_feat_compute_batched: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  B2K_CUDA_CHECK(cudaEventSynchronize(f->staging_free));
  int max_frames = 0;
  for (int i = 0; i < num_lanes; i++) {
    FeatLane &L = f->h_lanes[i];
    L.wave = d_wave ? d_wave[i] : nullptr; L.wave16 = d_wave16 ? d_wave16[i] : nullptr; L.num_samples = num_samples[i];
    if (!L.wave && !L.wave16 && num_frames[i] > 0) return set_error(B2K_ERR_INVALID, "b2k_feat_compute_batched: null waveform");
    L.first_frame = first_frame ? first_frame[i] : 0; L.num_frames = num_frames[i];
    L.out = d_out[i]; L.row_stride = row_stride;
    if (L.num_frames < 0 || (L.num_frames > 0 && L.num_samples <= 0))
      return set_error(B2K_ERR_INVALID, "b2k_feat_compute_batched: bad lane descriptor");
    // every requested frame must be computable from the samples present
    if (L.num_frames > 0 && L.first_frame + L.num_frames > b2k_feat_num_frames(f, L.num_samples, 1))
      return set_error(B2K_ERR_INVALID, "b2k_feat_compute_batched: frames beyond the available samples");
    if (L.num_frames > max_frames) max_frames = L.num_frames;
  }
  if (max_frames == 0) return B2K_OK;
  B2K_CUDA_CHECK(cudaMemcpyAsync(f->d_lanes, f->h_lanes, sizeof(FeatLane) * num_lanes, cudaMemcpyHostToDevice, st));
  B2K_CUDA_CHECK(cudaEventRecord(f->staging_free, st));
  const int frames_per_cta = FEAT_FRAMES_PER_CTA;
  dim3 grid((max_frames + frames_per_cta - 1) / frames_per_cta, num_lanes);
  feat_kernel<<<grid, FEAT_WARPS * 32, f->smem_bytes, st>>>(f->p, f->d_lanes, frames_per_cta, (int)f->stage_bytes);
  B2K_LAUNCH_CHECK();
  return B2K_OK;
}

int b2k_cmvn_apply_batched(b2k_feat *f, const b2k_cmvn_cfg *cfg, int32_t num_lanes,
                           const float *const *d_in, float *const *d_out, int32_t in_stride,
                           int32_t out_stride, const int32_t *first_frame, const int32_t *num_frames,
                           double *const *d_state, const double *d_global_stats,
                           const double *d_speaker_stats, void *stream) {
  if (!f || !cfg || num_lanes <= 0 || num_lanes > f->max_lanes || !d_in || !d_out || !num_frames || !d_state)
    return set_error(B2K_ERR_INVALID, "b2k_cmvn_apply_batched: bad args");
  if (!(cfg->speaker_frames <= cfg->cmn_window && cfg->global_frames <= cfg->speaker_frames))   // OnlineCmvnOptions::Check
    return set_error(B2K_ERR_INVALID, "OnlineCmvnOptions::Check failed");
  if (cfg->normalize_variance && !cfg->normalize_mean)
    return set_error(B2K_ERR_INVALID, "normalize_variance requires normalize_mean");
  if (!d_global_stats)
    return set_error(B2K_ERR_INVALID, "Global CMN stats are required");                         // online-feature.cc:417
  if (f->p.dim > 1024) return set_error(B2K_ERR_INVALID, "dim too large");
  cudaStream_t st = (cudaStream_t)stream;
  B2K_CUDA_CHECK(cudaEventSynchronize(f->staging_free));
  for (int i = 0; i < num_lanes; i++) {
    CmvnLane &L = f->h_clanes[i];
    L.in = d_in[i]; L.out = d_out[i]; L.in_stride = in_stride; L.out_stride = out_stride;
    L.first_frame = first_frame ? first_frame[i] : 0; L.num_frames = num_frames[i]; L.state = d_state[i];
    L.speaker = nullptr; L.speaker_out = nullptr; L.max_remembered_frames = -1.0f;
  }
  B2K_CUDA_CHECK(cudaMemcpyAsync(f->d_clanes, f->h_clanes, sizeof(CmvnLane) * num_lanes, cudaMemcpyHostToDevice, st));
  B2K_CUDA_CHECK(cudaEventRecord(f->staging_free, st));
  CmvnParams cp;
  cp.dim = f->p.dim; cp.cmn_window = cfg->cmn_window; cp.speaker_frames = cfg->speaker_frames;
  cp.global_frames = cfg->global_frames; cp.normalize_mean = cfg->normalize_mean;
  cp.normalize_variance = cfg->normalize_variance; cp.global_stats = d_global_stats; cp.speaker_stats = d_speaker_stats;
  int threads = ((cp.dim + 31) / 32) * 32;
  cmvn_kernel<<<num_lanes, threads, 0, st>>>(cp, f->d_clanes);
  B2K_LAUNCH_CHECK();
  return B2K_OK;
}

}  // extern "C"
