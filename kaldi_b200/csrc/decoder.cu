// decoder.cu — B200-native batched WFST lattice decoder (sm_100a).
//
// What it computes: the token-passing search of LatticeFasterDecoderTpl
// (reference decoder/lattice-faster-decoder.cc: InitDecoding :63, GetCutoff
// :653, ProcessEmitting :723, ProcessNonemitting :830, FinalizeDecoding :634,
// PruneForwardLinks[Final] :308/:385, GetRawLattice :114) behind the surface of
// cuda_decoder::CudaDecoder (cudadecoder/cuda-decoder.h:171-346).  It is NOT a
// port of cudadecoder/cuda-decoder-kernels.cu: that design launches ~25
// kernels and a blocking D2H per frame and dedups arc-instantiated tokens; this
// one keeps the CPU decoder's semantics (one token per (frame,state), all
// admitted links kept, float association (tok + (offset - ll)) + w, per-frame
// cost_offset) so that finalized lattices are comparable bit-for-bit with the
// CPU decoder (see oracle/decoder_oracle.cc, mode "order free").
//
// Execution model: ONE persistent CTA per lane (utterance).  Utterances are
// independent, so no grid-wide synchronisation exists anywhere: the CTA loops
// over all frames of its lane with only __syncthreads() between phases, reads
// the CSR HCLG and the log-likelihood rows from HBM/L2, and appends tokens and
// forward links to that channel's arenas in HBM.  Per frame:
//   1. best/cutoff   block reductions (+ exact radix-select for max/min-active)
//   2. seed pass     best token's arcs, association of :762-763
//   3. expand        thread-per-arc over a degree prefix sum (load balanced),
//                    min-reduction of tot_cost -> FINAL next_cutoff; arcs under
//                    a running upper bound are staged as candidates
//   4. admit         candidates with tot < final cutoff: hash insert (CAS),
//                    atomicMin on the token cost, forward link written (16 B)
//   5. epsilon       worklist relaxation to the fixpoint, then link generation
//   6. commit        costs -> arena, link dst slot -> token index, hash cleared
// Finalisation (backward pruning sweep with final costs) and lattice extraction
// are separate kernels with the same one-CTA-per-channel shape.
//
// Roofline class: HBM (irregular gather/scatter).  Algorithmic bytes per frame
// (DESIGN.md): 16 B per arc examined + 16 B per source token + 36 B per link
// admitted + 16 B per token kept.

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace b2k {

// ------------------------------------------------------------------ data types

struct FstDev {
  int32_t num_states, start;
  int32_t num_e, num_ne;
  const int2 *st_off;       // [N+1] {emitting offset, epsilon offset}
  const int4 *e_arcs;       // {nextstate, weight bits, pdf, olabel}
  const int4 *ne_arcs;      // {nextstate, weight bits, olabel | LAST<<31, eps offset of nextstate or -1}
  const int32_t *e_ilabel;  // [num_e] transition-id
  const float *final_cost;  // [N]
};

struct ChanState {          // per channel, device resident
  int32_t status;
  int32_t frames_decoded;   // -1 = InitDecoding not run
  int32_t ntok, nlink;      // arena fill
  int32_t finalized;
  int32_t lat_states, lat_arcs, lat_finals;
  int32_t any_final;
  float final_best_cost;
  int32_t hc;               // reference-order mode: HashList size (hash-list-inl.h:38), starts at 1000
  int32_t err_line;         // decoder.cu line that raised the first error of this channel (diagnostics)
  unsigned long long arcs_e, arcs_ne;
  unsigned long long prof[16];  // cycles per phase of the reference-order kernel (bench.py: decoder_phase_share)
};

#define B2K_SET_ERR(S, code) do { if (atomicCAS(&(S).err, 0, (code)) == 0) (S).err_line = __LINE__; } while (0)
#define B2K_EPS_FLAG 0x80000000u
#define B2K_ARC_MASK 0x3fffffffu
#define B2K_HASH_EMPTY (-1)

struct DecParams {
  FstDev fst;
  // config
  float beam, lattice_beam, beam_delta;
  int32_t max_active, min_active;
  int32_t max_tpf;          // tokens per frame capacity
  int32_t hash_size, hash_log;
  int32_t cand_cap;
  int32_t max_frames;
  int32_t max_tokens, max_links;
  // channel storage (arrays of [nchannels * capacity])
  ChanState *chan;
  int32_t *tok_state;
  float *tok_cost;
  float *tok_extra;
  int4 *links;              // {src tok, dst tok, arc id (bit31 = eps), acoustic bits}
  int32_t *frame_tok_begin; // [nch * (max_frames+2)]
  int32_t *frame_link_begin;// [nch * (max_frames+2)]
  int32_t *frame_link_eps;  // [nch * (max_frames+2)]
  float *frame_cost_offset; // [nch * (max_frames+1)]
  float *frame_cutoff;      // [nch * (max_frames+1)]
  // CTA scratch (arrays of [nslots * capacity]; a slot = one resident CTA of the persistent launch)
  int4 *hash;               // {key state, cost ord, tok idx in frame, stamp}
  int32_t *tokslot;         // [max_tpf]
  int32_t *wl;              // [2 * max_tpf]
  int32_t *cand;            // [5 * cand_cap] src, arc, next, tot bits, ac bits
  uint32_t *new_extra;      // [max_tpf] (finalize)
  int32_t *lane_stamp;      // [nslots]
  // compact lattice of a finalized channel (written by dec_finalize_kernel)
  int4 *lat_states;         // [nch * cap_ls] {frame, hclg state, tot bits, extra bits}
  int4 *lat_arcs;           // [nch * cap_la] {src id, dst id, ilabel, olabel}
  float2 *lat_arcw;         // [nch * cap_la] {graph cost, acoustic cost - cost_offset}
  int2 *lat_finals;         // [nch * cap_lf] {state id, final cost bits}
  int32_t cap_ls, cap_la, cap_lf;
  // reference-order mode scratch (per lane)
  uint32_t *x_bm;           // [pos_cap/32] bitmap of first-admission positions
  int32_t *x_wbase;         // [pos_cap/32]
  int32_t *x_by_ins;        // [max_tpf] insertion index -> hash slot
  int4 *x_bk;               // [hc_cap] per HashList bucket {first insertion index, population, fill cursor, -}
  int32_t *x_sbase;         // [max_tpf]
  int32_t *x_run;           // [max_tpf]
  int32_t *x_order;         // [max_tpf] list rank -> insertion index
  int32_t *x_xb;            // [max_tpf] insertion index -> HashList bucket (state % hash size)
  int4 *x_rec;              // [max_tpf] per dense token index {record offset, eps out-degree, #arcs admitted, replay cost bits}
  int32_t *x_newseq;        // [max_tpf] creation order found by the replay (eps-created tokens)
  int4 *x_adj;              // [adj_cap] eps arc entries {dest dense index, weight bits (+inf = not admitted), arc id, dest hash slot}
  int32_t *x_adjo;          // [adj_cap] owner (source dense index) of each entry
  int32_t adj_cap;
  int32_t rs_rcap, rs_ecap, rs_qcap;   // shared-memory replay capacities (tokens, arcs, worklist); 0 = off
  int32_t pos_cap, hc_cap, queue_cap;
  float hash_ratio;
  // per launch
  const int32_t *lane_channel;
  const float *const *lane_loglikes;
  const int32_t *lane_nframes;
  int32_t row_stride;
  int32_t do_init;
  // persistent launch: a CTA serves lanes until the counter runs out; its scratch is indexed by blockIdx.x
  int32_t n_lanes;
  int32_t *lane_counter;
  int32_t num_pdfs;         // log-likelihood columns the graph reads
  int32_t ll_smem;          // 1 = the frame's log-likelihood row is staged in shared memory (after the replay arrays)
  int32_t rs_bytes;         // bytes of the replay arrays at the start of dynamic shared memory
  // second-generation reference-order frame step (per-CTA scratch; see struct X2)
  int4 *v2_trec, *v2_tok4;
  uint32_t *v2_c0, *v2_adjc, *v2_qstamp, *v2_hw;
  int32_t *v2_pf;
  int32_t v2_hw_len;
  int32_t fin_scap;         // finalize: tokens of a list whose sweep state is held in shared memory (0 = off)
  int32_t hash_prefetch;    // second generation: L2 prefetch of the insert's first probe slot (experiment)
  int32_t par_walk;         // second generation: replay by connected components (one thread each)
  int32_t cid_smem;         // second generation: 16-bit compact ids of the replay in the upper half of the shared arc area
  int32_t v2_l1_shift;      // level-1 window = 2^shift x the reference's HashList size (rounded up to a power of two)
};

// ------------------------------------------------------------------ block helpers

template <int T>
__device__ __forceinline__ unsigned long long block_min_u64(unsigned long long v,
                                                            unsigned long long *sh) {
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w < v ? w : v;
  }
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  unsigned long long r = sh[0];
#pragma unroll
  for (int i = 1; i < T / 32; i++) r = sh[i] < r ? sh[i] : r;
  return r;
}

template <int T>
__device__ __forceinline__ uint32_t block_min_u32(uint32_t v, uint32_t *sh) {
  v = __reduce_min_sync(0xffffffffu, v);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  uint32_t r = sh[0];
#pragma unroll
  for (int i = 1; i < T / 32; i++) r = min(r, sh[i]);
  return r;
}

template <int T>
__device__ __forceinline__ int block_sum_i32(int v, int *sh) {
  v = __reduce_add_sync(0xffffffffu, v);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  int r = 0;
#pragma unroll
  for (int i = 0; i < T / 32; i++) r += sh[i];
  return r;
}

// exclusive scan of one int per thread; returns exclusive prefix, *total = sum
template <int T>
__device__ __forceinline__ int block_excl_scan(int v, int *sh /*[T/32+1]*/, int *total) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  __syncthreads();
  if (lane == 31) sh[warp] = incl;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < T / 32; i++) {
    int s = sh[i];
    if (i < warp) base += s;
    tot += s;
  }
  *total = tot;
  return base + incl - v;
}

// exact k-th smallest (0-indexed) of n float costs: 4-pass 8-bit radix select
template <int T>
__device__ float block_select_kth(const float *vals, int n, int k, uint32_t *hist /*[256]*/,
                                  uint32_t *sh_pref /*[2]*/) {
  uint32_t prefix = 0, mask = 0;
  int krem = k;
  for (int pass = 3; pass >= 0; pass--) {
    int shift = pass * 8;
    for (int i = threadIdx.x; i < 256; i += T) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += T) {
      uint32_t key = f2ord(vals[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int cum = 0, b = 0;
      for (; b < 256; b++) {
        int c = (int)hist[b];
        if (cum + c > krem) break;
        cum += c;
      }
      if (b > 255) b = 255;
      sh_pref[0] = (uint32_t)b;
      sh_pref[1] = (uint32_t)(krem - cum);
    }
    __syncthreads();
    prefix |= sh_pref[0] << shift;
    mask |= 255u << shift;
    krem = (int)sh_pref[1];
    __syncthreads();
  }
  return ord2f(prefix);
}

// phase timers (thread 0 only): TICK(s, i) adds the cycles since the previous tick to slot i
#define B2K_TICK(S, I) do { if (PROF && threadIdx.x == 0) { long long _t = clock64(); (S).prof[(I)] += (unsigned long long)(_t - (S).tlast); (S).tlast = _t; } } while (0)

// ------------------------------------------------------------------ hash

struct LaneCtx {
  int4 *hash;
  int32_t *tokslot;
  int32_t *tok_state;   // channel arena
  int32_t tbase;
  int32_t hash_mask, hash_log, max_tpf, max_tokens;
  int32_t l1_mask, l1_log;   // level-1 window of the table for this frame (see probe_slot)
  int *used_l2;              // shared: set when a bucket-keyed insert went past its level-1 window this frame
  uint32_t hc;               // the reference's HashList size of this frame and its fastmod constant (bucket_b)
  unsigned long long hc_m;
  int *ntok_new;        // shared
  int *err;             // shared
  int *err_line;        // shared
};

// Two-level open addressing.  A CTA's table has hash_size slots (capacity), but a frame only
// ever holds a few thousand tokens: the first B2K_L1_PROBES probes of a state stay inside the
// first l1_size slots (l1_size ~ 3x the previous frame's token count, chosen per frame), so the
// slots a frame touches are a small, L2-resident prefix of the table; only a state whose whole
// level-1 window is occupied goes on to linear probing over the full table with a second hash.
// Slots are never freed inside a frame, so a state's position is stable: if it is in its
// level-1 window every later lookup finds it there, and if it is not, the window was full when
// it was inserted and is still full, so later lookups also fall through to level 2.
#define B2K_L1_PROBES 16

__device__ __forceinline__ uint32_t probe_slot(const LaneCtx &c, int32_t state, int i) {
  if (i < B2K_L1_PROBES) return ((((uint32_t)state * 2654435761u) >> (32 - c.l1_log)) + (uint32_t)i) & (uint32_t)c.l1_mask;
  return ((((uint32_t)state * 0x85ebca6bu) >> (32 - c.hash_log)) + (uint32_t)(i - B2K_L1_PROBES)) & (uint32_t)c.hash_mask;
}

// level-1 size for a frame.  Measured (profiles/r02_decoder_history.md): a level-1 window sized to the frame made probe
// sequences longer and did not make the table L2 resident (a resident CTA touches ~0.85 MB of scratch per frame, 296 of
// them 2x the L2), so the first-generation kernels use the whole table as level 1; the bucket-keyed table of the
// second-generation frame step (probe_slot_b) sizes its window from the reference's own HashList size.
__device__ __forceinline__ void set_l1(LaneCtx &c, int K) {
  (void)K;
  c.l1_log = c.hash_log;
  c.l1_mask = c.hash_mask;
}

// find-or-insert; returns slot (or -1 after an overflow was flagged)
__device__ __forceinline__ int hash_insert(const LaneCtx &c, int32_t state) {
  for (int probe = 0; probe <= c.hash_mask + B2K_L1_PROBES; probe++) {
    const uint32_t h = probe_slot(c, state, probe);
    int *keyp = reinterpret_cast<int *>(&c.hash[h]);
    int old = atomicCAS(keyp, B2K_HASH_EMPTY, state);
    if (old == B2K_HASH_EMPTY) {
      int idx = atomicAdd(c.ntok_new, 1);
      if (idx < c.max_tpf && c.tbase + idx < c.max_tokens) {
        c.tok_state[c.tbase + idx] = state;
        c.tokslot[idx] = (int)h;
        keyp[2] = idx;
      } else {
        do { if (atomicCAS(c.err, 0, B2K_ERR_OVERFLOW) == 0) *c.err_line = __LINE__; } while (0);
      }
      return (int)h;
    }
    if (old == state) return (int)h;
  }
  do { if (atomicCAS(c.err, 0, B2K_ERR_OVERFLOW) == 0) *c.err_line = __LINE__; } while (0);
  return -1;
}

__device__ __forceinline__ int hash_find(const LaneCtx &c, int32_t state) {
  for (int probe = 0; probe <= c.hash_mask + B2K_L1_PROBES; probe++) {
    const uint32_t h = probe_slot(c, state, probe);
    int key = *reinterpret_cast<volatile int *>(&c.hash[h]);
    if (key == state) return (int)h;
    if (key == B2K_HASH_EMPTY) return -1;
  }
  return -1;
}

template <int T>
__device__ void reset_lane_hash(int4 *hash, int hash_size) {
  __syncthreads();
  for (int i = threadIdx.x; i < hash_size; i += T) {
    hash[i].x = B2K_HASH_EMPTY;
    hash[i].y = (int)B2K_INF_ORD;
    hash[i].z = 0x7fffffff;   // reference-order mode keeps the insertion seq here
  }
}

// ------------------------------------------------------------------ the frame kernel

template <int T>
struct __align__(16) DecShared {
  unsigned long long red64[T / 32];
  uint32_t red32[T / 32];
  int redi[T / 32 + 1];
  uint32_t hist[256];
  uint32_t pref[2];
  static constexpr int TPT = (T >= 512) ? 1024 / T : 4;   // tokens per thread in an expansion chunk (shared memory is
                                                          // also L1: 2048-token chunks cost more in hit rate than they save)
  static constexpr int CT = T * TPT;                      // tokens per expansion chunk
  int chunk_off[CT + 1];
  int chunk_ebeg[CT];
  float chunk_cost[CT];
  int chunk_d[T];
  int ntok_new, nlink_new, ncand, err, err_line;
  int wl_n[2];
  uint32_t running_ord;
  int stamp;
  int cont;
  float scanf_[2][T / 32];
  int q_n;
  int used_l2;
  int rs_n, rs_e, rs_ok, rs_eov;
  unsigned long long prof[16];
  long long tlast;
};

// epsilon closure + link generation + commit of the frame being built.
// On entry: the emitting phase (or init) has inserted tokens into the hash.
template <int T>
__device__ void finish_frame(const DecParams &p, DecShared<T> &s, const LaneCtx &ctx, int lane,
                             int ch, int list_index, float cutoff, float cost_offset,
                             int32_t lbase, unsigned long long *arcs_ne_acc) {
  const int tid = threadIdx.x;
  const FstDev &g = p.fst;
  int4 *hash = ctx.hash;
  int32_t *wl0 = p.wl + (size_t)lane * 2 * p.max_tpf;
  int32_t *wl1 = wl0 + p.max_tpf;
  int4 *links = p.links + (size_t)ch * p.max_links;
  float *tok_cost = p.tok_cost + (size_t)ch * p.max_tokens;

  __syncthreads();
  const int n_emit_links = s.nlink_new;
  const int N1 = min(s.ntok_new, p.max_tpf);
  // initial worklist = every token created so far (eps-degree checked at pop)
  for (int i = tid; i < N1; i += T) wl0[i] = ctx.tokslot[i];
  if (tid == 0) { s.wl_n[0] = N1; s.wl_n[1] = 0; s.cont = (N1 > 0 && !s.err); }
  __syncthreads();
  int cur = 0;
  unsigned long long ne_count = 0;
  while (s.cont) {
    const int n = s.wl_n[cur];
    const int stamp = s.stamp + 1;
    int32_t *in = cur ? wl1 : wl0;
    int32_t *out = cur ? wl0 : wl1;
    for (int k = tid; k < n; k += T) {
      int slot = in[k];
      int4 *sp = &hash[slot];
      int state = *reinterpret_cast<volatile int *>(&sp->x);
      float c = ord2f(*reinterpret_cast<volatile uint32_t *>(&sp->y));
      if (!(c < cutoff)) continue;
      int2 o0 = __ldg(&g.st_off[state]), o1 = __ldg(&g.st_off[state + 1]);
      ne_count += (unsigned long long)(o1.y - o0.y);
      for (int a = o0.y; a < o1.y; a++) {
        int4 arc = __ldg(&g.ne_arcs[a]);
        float tot = c + __int_as_float(arc.y);
        if (tot < cutoff) {
          int ds = hash_insert(ctx, arc.x);
          if (ds < 0) break;
          uint32_t nv = f2ord(tot);
          uint32_t old = atomicMin(reinterpret_cast<uint32_t *>(&hash[ds].y), nv);
          if (nv < old) {
            if (atomicExch(&hash[ds].w, stamp) != stamp) {
              int q = atomicAdd(&s.wl_n[cur ^ 1], 1);
              if (q < p.max_tpf) out[q] = ds;
              else B2K_SET_ERR(s, B2K_ERR_OVERFLOW);
            }
          }
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      s.wl_n[cur] = 0;
      s.stamp = stamp;
      s.cont = (min(s.wl_n[cur ^ 1], p.max_tpf) > 0 && !s.err);
      if (s.wl_n[cur ^ 1] > p.max_tpf) s.wl_n[cur ^ 1] = p.max_tpf;
    }
    cur ^= 1;
    __syncthreads();
  }
  __syncthreads();
  // epsilon links: {tok, eps arc} with final cost + w < cutoff
  const int Nall = min(s.ntok_new, p.max_tpf);
  const int32_t *tok_state = ctx.tok_state;
  for (int i = tid; i < Nall; i += T) {
    int slot = ctx.tokslot[i];
    float c = ord2f((uint32_t)hash[slot].y);
    if (!(c < cutoff)) continue;
    int state = tok_state[ctx.tbase + i];
    int2 o0 = __ldg(&g.st_off[state]), o1 = __ldg(&g.st_off[state + 1]);
    for (int a = o0.y; a < o1.y; a++) {
      int4 arc = __ldg(&g.ne_arcs[a]);
      float tot = c + __int_as_float(arc.y);
      if (tot < cutoff) {
        int ds = hash_find(ctx, arc.x);
        int li = atomicAdd(&s.nlink_new, 1);
        if (lbase + li < p.max_links && ds >= 0)
          links[lbase + li] = make_int4(ctx.tbase + i, ds, (int)((uint32_t)a | B2K_EPS_FLAG), 0);
        else
          B2K_SET_ERR(s, B2K_ERR_OVERFLOW);
      }
    }
  }
  __syncthreads();
  // commit: costs to arena, link dst slot -> arena token index
  const int nlink = s.nlink_new;
  for (int i = tid; i < Nall; i += T) {
    int slot = ctx.tokslot[i];
    tok_cost[ctx.tbase + i] = ord2f((uint32_t)hash[slot].y);
  }
  if (!s.err) {
    for (int li = tid; li < nlink; li += T) {
      int4 *lp = &links[lbase + li];
      int slot = lp->y;
      lp->y = ctx.tbase + hash[slot].z;
    }
  }
  __syncthreads();
  for (int i = tid; i < Nall; i += T) {
    int slot = ctx.tokslot[i];
    hash[slot].x = B2K_HASH_EMPTY;
    hash[slot].y = (int)B2K_INF_ORD;
  }
  if (tid == 0) {
    size_t fo = (size_t)ch * (p.max_frames + 2);
    p.frame_tok_begin[fo + list_index] = ctx.tbase;
    p.frame_tok_begin[fo + list_index + 1] = ctx.tbase + Nall;
    p.frame_link_begin[fo + list_index] = lbase;
    p.frame_link_eps[fo + list_index] = lbase + n_emit_links;
    p.frame_link_begin[fo + list_index + 1] = lbase + nlink;
    if (list_index > 0) {
      size_t co = (size_t)ch * (p.max_frames + 1) + (list_index - 1);
      p.frame_cost_offset[co] = cost_offset;
      p.frame_cutoff[co] = cutoff;
    }
  }
  // warp-reduce eps arc counter into the caller's accumulator (thread 0 adds)
  for (int o = 16; o > 0; o >>= 1) ne_count += __shfl_xor_sync(0xffffffffu, ne_count, o);
  if ((tid & 31) == 0 && ne_count) atomicAdd(arcs_ne_acc, ne_count);
  __syncthreads();
}

template <int T>
__device__ void dec_advance_lane(const DecParams &p, DecShared<T> &s, const int lane, const int slot) {
  const int tid = threadIdx.x;
  const int ch = p.lane_channel[lane];
  ChanState *cs = &p.chan[ch];
  const FstDev &g = p.fst;
  const float kInf = __int_as_float(0x7f800000);

  if (cs->status != B2K_OK) return;
  if (!p.do_init && cs->frames_decoded < 0) {
    if (tid == 0) { cs->status = B2K_ERR_STATE; cs->err_line = __LINE__; }
    return;
  }

  int32_t *tok_state = p.tok_state + (size_t)ch * p.max_tokens;
  float *tok_cost = p.tok_cost + (size_t)ch * p.max_tokens;
  int4 *links = p.links + (size_t)ch * p.max_links;
  int32_t *cand = p.cand + (size_t)slot * 5 * p.cand_cap;
  int32_t *c_src = cand, *c_arc = cand + p.cand_cap, *c_next = cand + 2 * (size_t)p.cand_cap,
          *c_tot = cand + 3 * (size_t)p.cand_cap, *c_ac = cand + 4 * (size_t)p.cand_cap;

  LaneCtx ctx;
  ctx.hash = p.hash + (size_t)slot * p.hash_size;
  ctx.tokslot = p.tokslot + (size_t)slot * p.max_tpf;
  ctx.tok_state = tok_state;
  ctx.hash_mask = p.hash_size - 1;
  ctx.hash_log = p.hash_log;
  ctx.max_tpf = p.max_tpf;
  ctx.max_tokens = p.max_tokens;
  ctx.ntok_new = &s.ntok_new;
  ctx.err = &s.err;
  ctx.err_line = &s.err_line;
  set_l1(ctx, 0);

  if (tid == 0) { s.err = 0; s.err_line = 0; s.stamp = p.lane_stamp[slot]; }
  __syncthreads();

  if (p.do_init) {
    // InitDecoding: start token, cost 0, then ProcessNonemitting(config.beam)
    if (tid == 0) { s.ntok_new = 0; s.nlink_new = 0; }
    __syncthreads();
    ctx.tbase = 0;
    if (tid == 0) {
      int slot = hash_insert(ctx, g.start);
      if (slot >= 0) atomicMin(reinterpret_cast<uint32_t *>(&ctx.hash[slot].y), f2ord(0.0f));
    }
    finish_frame<T>(p, s, ctx, slot, ch, 0, p.beam, 0.0f, 0, &cs->arcs_ne);
    if (tid == 0) {
      cs->frames_decoded = 0;
      cs->ntok = min(s.ntok_new, p.max_tpf);
      cs->nlink = s.nlink_new;
      cs->finalized = 0;
      if (s.err) { cs->status = s.err; cs->err_line = s.err_line; }
      p.lane_stamp[slot] = s.stamp;
    }
    if (s.err) reset_lane_hash<T>(ctx.hash, p.hash_size);
    return;
  }

  const int nframes = p.lane_nframes[lane];
  const float *ll_base = p.lane_loglikes[lane];
  int frames_decoded = cs->frames_decoded;
  int32_t tbase = cs->ntok, lbase = cs->nlink;
  unsigned long long arcs_e_total = 0;   // thread 0 accumulates per-chunk totals
  const size_t fo = (size_t)ch * (p.max_frames + 2);

  for (int fi = 0; fi < nframes; fi++) {
    if (frames_decoded >= p.max_frames) { if (tid == 0) B2K_SET_ERR(s, B2K_ERR_OVERFLOW); __syncthreads(); break; }
    const float *ll = ll_base + (size_t)fi * p.row_stride;
    const int pb = p.frame_tok_begin[fo + frames_decoded];
    const int pe = p.frame_tok_begin[fo + frames_decoded + 1];
    const int K = pe - pb;
    set_l1(ctx, K);

    // ---- 1. best token and cutoff (GetCutoff :653-720)
    unsigned long long local = ~0ull;
    for (int i = pb + tid; i < pe; i += T) {
      unsigned long long key = ((unsigned long long)f2ord(tok_cost[i]) << 32) | (uint32_t)tok_state[i];
      local = key < local ? key : local;
    }
    unsigned long long bestkey = block_min_u64<T>(local, s.red64);
    float best_cost = kInf;
    int best_state = -1;
    if (K > 0) { best_cost = ord2f((uint32_t)(bestkey >> 32)); best_state = (int)(uint32_t)(bestkey & 0xffffffffu); }
    const float beam_cutoff = best_cost + p.beam;
    float cur_cutoff = beam_cutoff, adaptive_beam = p.beam;
    if (K > 0 && !(p.max_active == 0x7fffffff && p.min_active == 0)) {
      int c_lt = 0, c_le = 0;
      for (int i = pb + tid; i < pe; i += T) {
        float c = tok_cost[i];
        c_lt += (c < beam_cutoff);
        c_le += (c <= beam_cutoff);
      }
      c_lt = block_sum_i32<T>(c_lt, s.redi);
      c_le = block_sum_i32<T>(c_le, s.redi);
      bool done = false;
      if (K > p.max_active && c_lt > p.max_active) {
        // max_active_cutoff < beam_cutoff  (:689-699)
        float mac = block_select_kth<T>(tok_cost + pb, K, p.max_active, s.hist, s.pref);
        cur_cutoff = mac;
        adaptive_beam = mac - best_cost + p.beam_delta;
        done = true;
      }
      if (!done) {
        float min_active_cutoff = kInf;
        if (K > p.min_active) {
          if (p.min_active == 0) min_active_cutoff = best_cost;
          else if (c_le <= p.min_active)
            min_active_cutoff = block_select_kth<T>(tok_cost + pb, K, p.min_active, s.hist, s.pref);
          else min_active_cutoff = beam_cutoff;  // known to be <= beam_cutoff
        }
        if (min_active_cutoff > beam_cutoff) {   // :711-714
          adaptive_beam = min_active_cutoff - best_cost + p.beam_delta;
          cur_cutoff = min_active_cutoff;
        }
      }
    }
    const float cost_offset = (K > 0) ? -best_cost : 0.0f;

    // ---- 2. seed next_cutoff from the best token (:753-768)
    uint32_t seed_local = B2K_INF_ORD;
    if (K > 0) {
      int2 o0 = __ldg(&g.st_off[best_state]), o1 = __ldg(&g.st_off[best_state + 1]);
      for (int a = o0.x + tid; a < o1.x; a += T) {
        int4 arc = __ldg(&g.e_arcs[a]);
        float new_weight = __int_as_float(arc.y) + cost_offset - __ldg(&ll[arc.z]) + best_cost;
        seed_local = min(seed_local, f2ord(new_weight + adaptive_beam));
      }
    }
    const float seed_cutoff = ord2f(block_min_u32<T>(seed_local, s.red32));

    // ---- 3. expand (ProcessEmitting main loop :779-812), pass A
    if (tid == 0) { s.ntok_new = 0; s.nlink_new = 0; s.ncand = 0; s.running_ord = B2K_INF_ORD; }
    __syncthreads();
    for (int cb = pb; cb < pe; cb += T) {
      int i = cb + tid;
      int deg = 0, ebeg = 0;
      float c = 0.f;
      if (i < pe) {
        c = tok_cost[i];
        if (c <= cur_cutoff) {
          int st = tok_state[i];
          int2 o0 = __ldg(&g.st_off[st]), o1 = __ldg(&g.st_off[st + 1]);
          ebeg = o0.x;
          deg = o1.x - o0.x;
        }
      }
      int total;
      int off = block_excl_scan<T>(deg, s.redi, &total);
      s.chunk_off[tid] = off;
      s.chunk_ebeg[tid] = ebeg;
      s.chunk_cost[tid] = c;
      if (tid == 0) { s.chunk_off[T] = total; arcs_e_total += (unsigned long long)total; }
      __syncthreads();
      const int rounds = (total + T - 1) / T;
      for (int r = 0; r < rounds; r++) {
        int j = r * T + tid;
        bool active = j < total;
        float tot = kInf, ac = 0.f;
        int src = 0, a = 0, nexts = 0;
        if (active) {
          // largest t with chunk_off[t] <= j
          int lo = 0, hi = T;
          while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (s.chunk_off[mid] <= j) lo = mid; else hi = mid;
          }
          a = s.chunk_ebeg[lo] + (j - s.chunk_off[lo]);
          src = cb + lo;
          int4 arc = __ldg(&g.e_arcs[a]);
          ac = cost_offset - __ldg(&ll[arc.z]);
          tot = s.chunk_cost[lo] + ac + __int_as_float(arc.y);
          nexts = arc.x;
        }
        uint32_t wmin = __reduce_min_sync(0xffffffffu, f2ord(tot));
        uint32_t run = *reinterpret_cast<volatile uint32_t *>(&s.running_ord);
        if (wmin < run) {
          if ((tid & 31) == 0) atomicMin(&s.running_ord, wmin);
          run = wmin;
        }
        float bound = fminf(seed_cutoff, ord2f(run) + adaptive_beam);
        bool is_cand = active && (tot < bound);
        uint32_t m = __ballot_sync(0xffffffffu, is_cand);
        if (m) {
          int lane_id = tid & 31;
          int base = 0;
          if (lane_id == 0) base = atomicAdd(&s.ncand, __popc(m));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (is_cand) {
            int idx = base + __popc(m & ((1u << lane_id) - 1u));
            if (idx < p.cand_cap) {
              c_src[idx] = src; c_arc[idx] = a; c_next[idx] = nexts;
              c_tot[idx] = __float_as_int(tot); c_ac[idx] = __float_as_int(ac);
            } else {
              B2K_SET_ERR(s, B2K_ERR_OVERFLOW);
            }
          }
        }
      }
      __syncthreads();
    }
    // FINAL next_cutoff = min(seed, min_tot + adaptive_beam)
    const float min_tot = ord2f(s.running_ord);
    const float next_cutoff = fminf(seed_cutoff, min_tot + adaptive_beam);
    const int ncand = min(s.ncand, p.cand_cap);

    // ---- 4. admit candidates (FindOrAddToken :261-302 + ForwardLink :804-806)
    ctx.tbase = tbase;
    for (int base_i = 0; base_i < ncand; base_i += T) {
      int idx = base_i + tid;
      bool adm = false;
      float tot = 0.f;
      if (idx < ncand) { tot = __int_as_float(c_tot[idx]); adm = tot < next_cutoff; }
      int slot = -1;
      if (adm) {
        slot = hash_insert(ctx, c_next[idx]);
        if (slot >= 0) atomicMin(reinterpret_cast<uint32_t *>(&ctx.hash[slot].y), f2ord(tot));
        else adm = false;
      }
      uint32_t m = __ballot_sync(0xffffffffu, adm);
      if (m) {
        int lane_id = tid & 31;
        int lb = 0;
        if (lane_id == 0) lb = atomicAdd(&s.nlink_new, __popc(m));
        lb = __shfl_sync(0xffffffffu, lb, 0);
        if (adm) {
          int li = lb + __popc(m & ((1u << lane_id) - 1u));
          if (lbase + li < p.max_links)
            links[lbase + li] = make_int4(c_src[idx], slot, c_arc[idx], c_ac[idx]);
          else
            B2K_SET_ERR(s, B2K_ERR_OVERFLOW);
        }
      }
    }
    // ---- 5/6. epsilon closure, eps links, commit
    finish_frame<T>(p, s, ctx, slot, ch, frames_decoded + 1, next_cutoff, cost_offset, lbase,
                    &cs->arcs_ne);
    if (s.err) break;
    tbase += min(s.ntok_new, p.max_tpf);
    lbase += s.nlink_new;
    frames_decoded++;
    __syncthreads();
  }
  if (tid == 0) {
    cs->frames_decoded = frames_decoded;
    cs->ntok = tbase;
    cs->nlink = lbase;
    cs->arcs_e += arcs_e_total;
    if (s.err) { cs->status = s.err; cs->err_line = s.err_line; }
    p.lane_stamp[slot] = s.stamp;
  }
  // an overflow can leave uncommitted tokens in the CTA's hash: wipe it so the
  // table is clean for the next lane it serves
  if (s.err) reset_lane_hash<T>(ctx.hash, p.hash_size);
}

// Persistent launch shared by the three per-lane kernels: the grid is at most the number of CTAs the
// device keeps resident, every CTA owns one scratch slot (hash table, per-frame arrays: indexed by
// blockIdx.x) and serves lanes until the launch's counter runs out.  The scratch a launch touches is
// therefore (resident CTAs) x (what a frame touches), which stays in L2, instead of growing with the batch.
#define B2K_PERSISTENT_LANES(BODY)                                            \
  __shared__ int s_lane_;                                                     \
  for (;;) {                                                                  \
    __syncthreads();                                                          \
    if (threadIdx.x == 0) s_lane_ = atomicAdd(p.lane_counter, 1);             \
    __syncthreads();                                                          \
    const int lane_ = s_lane_;                                                \
    if (lane_ >= p.n_lanes) break;                                            \
    BODY;                                                                     \
  }

template <int T>
__global__ void __launch_bounds__(T) dec_advance_kernel(DecParams p) {
  __shared__ DecShared<T> s;
  B2K_PERSISTENT_LANES(dec_advance_lane<T>(p, s, lane_, (int)blockIdx.x))
}


// ====================================================================== reference-order mode
//
// Bit-exact emulation of the order dependence of the CPU decoder (DESIGN.md
// "Why iteration order matters").  ProcessEmitting prunes against a RUNNING
// next_cutoff (lattice-faster-decoder.cc:794-796) while it walks the previous
// frame's tokens in HashList order; that running value is exactly an exclusive
// prefix-min over the (token, arc) sequence:
//     RC(p) = min(seed, min_{q<p} tot_q + adaptive_beam)
// (non-admitted arcs cannot lower it because adaptive_beam > 0), so admission
// is a block-wide scan, not a sequential walk -- PROVIDED each frame's tokens
// are stored in the arena in HashList iteration order.  That order is
// (first-occupancy rank of bucket state % hash_size, insertion order within
// the bucket) (hash-list-inl.h:126-175) and is rebuilt per frame from
// insertion sequence numbers: for tokens created by ProcessEmitting the
// sequence number is the position of the first admitted arc (atomicMin), for
// tokens created by ProcessNonemitting it comes from a literal, single-thread
// replay of the LIFO worklist (:858-896) -- the one inherently sequential piece.

struct XScratch {
  uint32_t *bm; int32_t *wbase, *by_ins, *sbase, *run, *order, *queue, *xb; int4 *bk;
  int4 *rec; int32_t *newseq; int4 *adj; int32_t *adjo;
};

// find-or-insert without arena write; *created tells whether this call made the token
__device__ __forceinline__ int hash_insert_x(const LaneCtx &c, int32_t state, bool *created, int *idx_out) {
  *created = false;
  for (int probe = 0; probe <= c.hash_mask + B2K_L1_PROBES; probe++) {
    const uint32_t h = probe_slot(c, state, probe);
    int *keyp = reinterpret_cast<int *>(&c.hash[h]);
    int old = atomicCAS(keyp, B2K_HASH_EMPTY, state);
    if (old == B2K_HASH_EMPTY) {
      int idx = atomicAdd(c.ntok_new, 1);
      if (idx < c.max_tpf) c.tokslot[idx] = (int)h;
      else do { if (atomicCAS(c.err, 0, B2K_ERR_OVERFLOW) == 0) *c.err_line = __LINE__; } while (0);
      *created = true;
      *idx_out = idx;
      return (int)h;
    }
    if (old == state) return (int)h;
  }
  do { if (atomicCAS(c.err, 0, B2K_ERR_OVERFLOW) == 0) *c.err_line = __LINE__; } while (0);
  return -1;
}

// HashList iteration order of the N tokens by_ins[0..N) (slot.z = insertion index):
// x.order[list rank] = insertion index.  Buckets are ranked by their first
// insertion (hash-list-inl.h:126-175); a bucket's tokens keep insertion order.
// Two steps: bucket_scatter() records, per bucket, its first insertion index and its
// population for the tokens [k0, k1) (it can be called again for later insertions:
// they never lower an existing bucket's first index); order_finish() turns that into
// the list order of all N tokens and clears the bucket arrays.
template <int T>
__device__ void bucket_scatter(int k0, int k1, int Hc, const int4 *hash, const XScratch &x) {
  __syncthreads();
  for (int k = k0 + threadIdx.x; k < k1; k += T) {
    int slot = x.by_ins[k];
    int b = (int)((uint32_t)hash[slot].x % (uint32_t)Hc);
    x.xb[k] = b;
    atomicMin(&x.bk[b].x, k);
    atomicAdd(&x.bk[b].y, 1);
  }
  __syncthreads();
}

template <int T>
__device__ void order_finish(int N, const XScratch &x, DecShared<T> &s, int *order_slot = nullptr) {
  const int tid = threadIdx.x;
  const int4 kEmptyBucket = make_int4(0x7fffffff, 0, 0, 0);
  __syncthreads();
  // list position of every bucket head = exclusive scan of the bucket populations in
  // first-insertion order.  A bucket with one token (the common case) is finished here.
  int carry = 0;
  for (int base = 0; base < N; base += T) {
    int k = base + tid, w = 0, b = 0;
    if (k < N) {
      b = x.xb[k];
      int4 bk = x.bk[b];
      if (bk.x == k) w = bk.y;
    }
    int total;
    int excl = block_excl_scan<T>(w, s.redi, &total);
    if (k < N) {
      if (w == 1) {
        x.order[carry + excl] = k;
        if (order_slot) order_slot[carry + excl] = x.by_ins[k];
        x.bk[b] = kEmptyBucket;
        x.xb[k] = -1;
      } else {
        x.sbase[k] = carry + excl;                   // only read through bucket heads
      }
    }
    carry += total;
  }
  __syncthreads();
  for (int k = tid; k < N; k += T) {
    int b = x.xb[k];
    if (b < 0) continue;
    int rb = x.sbase[x.bk[b].x];
    int q = atomicAdd(&x.bk[b].z, 1);
    x.run[rb + q] = k;
  }
  __syncthreads();
  for (int k = tid; k < N; k += T) {
    int b = x.xb[k];
    if (b < 0) continue;
    int4 bk = x.bk[b];
    int rb = x.sbase[bk.x];
    int within = 0;
    const int pop = bk.y;
    for (int j = 0; j < pop; j++) within += (x.run[rb + j] < k);
    x.order[rb + within] = k;
    if (order_slot) order_slot[rb + within] = x.by_ins[k];
  }
  __syncthreads();
  for (int k = tid; k < N; k += T) {
    int b = x.xb[k];
    if (b < 0) continue;
    x.bk[b] = kEmptyBucket;
  }
  __syncthreads();
}

// ProcessNonemitting replay + ordering + eps links + commit (reference order).
// On entry N1 = s.ntok_new tokens exist with slot.z = insertion index and
// by_ins filled for them.
template <int T, bool PROF>
__device__ void finish_frame_exact(const DecParams &p, DecShared<T> &s, const LaneCtx &ctx,
                                   const XScratch &x, int lane, int ch, int list_index, float cutoff,
                                   float cost_offset, int32_t lbase, int Hc, ChanState *cs) {
  const int tid = threadIdx.x;
  const FstDev &g = p.fst;
  int4 *hash = ctx.hash;
  int4 *links = p.links + (size_t)ch * p.max_links;
  int32_t *tok_state = p.tok_state + (size_t)ch * p.max_tokens;
  float *tok_cost = p.tok_cost + (size_t)ch * p.max_tokens;
  __syncthreads();
  if (s.err) return;     // uniform: nothing writes err between the barrier and this read
  const int n_emit_links = s.nlink_new;
  const int N1 = min(s.ntok_new, p.max_tpf);
  // bucket bookkeeping of the emitting tokens: the list order (which defines the initial
  // worklist, :852-856) is only materialised for the few tokens the replay needs
  bucket_scatter<T>(0, N1, Hc, hash, x);
  B2K_TICK(s, 3);
  // ---- eps closure.  The final costs and the token set are order independent
  // (least fixpoint), so they are computed by parallel relaxation; the literal
  // LIFO replay of ProcessNonemitting (:858-896) is then needed only for the
  // CREATION ORDER of the eps-created tokens.  The relaxation records, for
  // every token it expands, the eps arcs it admits; the LAST expansion of a
  // token runs at its final cost, so the last record is exactly
  // {arcs with final_cost + w < cutoff}: the arcs that become eps links, and a
  // superset of what the replay can admit (its cur_cost >= final cost).  The
  // replay then walks those dense records only.
  int32_t *wl0 = p.wl + (size_t)lane * 2 * p.max_tpf;
  int32_t *wl1 = wl0 + p.max_tpf;
  // cost after ProcessEmitting per dense index, and the first worklist: only the tokens
  // that are below the cutoff and have eps arcs at all (compacted, with what the first
  // round needs: {cost, first eps arc, eps out-degree, dense index})
  int4 *wlx = reinterpret_cast<int4 *>(x.queue);              // idle until the replay worklist is built
  if (tid == 0) { s.wl_n[0] = 0; s.wl_n[1] = 0; s.q_n = 0; }
  __syncthreads();
  for (int base = 0; base < N1; base += T) {
    const int d = base + tid;
    int deg = 0, ebeg = 0, cbits = 0;
    if (d < N1) {
      const int4 hs = hash[x.by_ins[d]];
      const float c0 = ord2f((uint32_t)hs.y);
      cbits = __float_as_int(c0);
      x.rec[d] = make_int4(0, 0, 0, cbits);
      if (c0 < cutoff) {
        int2 o0 = __ldg(&g.st_off[hs.x]), o1 = __ldg(&g.st_off[hs.x + 1]);
        ebeg = o0.y; deg = o1.y - o0.y;
      }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, deg > 0);
    if (m) {
      int wb = 0;
      if ((tid & 31) == 0) wb = atomicAdd(&s.wl_n[0], __popc(m));
      wb = __shfl_sync(0xffffffffu, wb, 0);
      if (deg > 0) wlx[wb + __popc(m & ((1u << (tid & 31)) - 1u))] = make_int4(cbits, ebeg, deg, d);
    }
  }
  __syncthreads();
  if (tid == 0) s.cont = (s.wl_n[0] > 0 && !s.err);
  __syncthreads();
  B2K_TICK(s, 4);
  // Each relaxation round expands its worklist ARC-parallel (eps out-degrees are very
  // skewed: word-boundary hubs carry thousands of eps arcs): a token's record is a
  // contiguous block of one entry per eps arc, in arc order; arcs that fail the
  // cutoff get weight +inf so that the replay skips them.
  {
    int cur = 0;
    bool first_round = true;
    const int kInfBits = 0x7f800000;
    while (s.cont) {
      const int n = s.wl_n[cur];
      const int stamp = s.stamp + 1;
      int32_t *in = cur ? wl1 : wl0;
      int32_t *out = cur ? wl0 : wl1;
      for (int cb = 0; cb < n; cb += T) {
        const int k = cb + tid;
        int deg = 0, ebeg = 0, dd = 0;
        float c = 0.f;
        if (k < n) {
          if (first_round) {
            // (a cost lowered since the snapshot re-queues the token, so the snapshot is as good as a fresh read)
            const int4 w = wlx[k];
            c = __int_as_float(w.x); ebeg = w.y; deg = w.z; dd = w.w;
          } else {
            int4 hs = __ldcg(&hash[in[k]]);                   // cost may be lowered concurrently: read at L2
            c = ord2f((uint32_t)hs.y);
            if (c < cutoff) {
              int2 o0 = __ldg(&g.st_off[hs.x]), o1 = __ldg(&g.st_off[hs.x + 1]);
              ebeg = o0.y; deg = o1.y - o0.y; dd = hs.z;
            }
          }
        }
        int total;
        const int off = block_excl_scan<T>(deg, s.redi, &total);
        s.chunk_off[tid] = off; s.chunk_ebeg[tid] = ebeg; s.chunk_cost[tid] = c; s.chunk_d[tid] = dd;
        if (tid == 0) {
          s.chunk_off[T] = total;
          s.ncand = s.q_n;                                    // record space of this chunk
          s.q_n += total;
          if (s.q_n > p.adj_cap) B2K_SET_ERR(s, B2K_ERR_OVERFLOW);
        }
        __syncthreads();
        if (s.err) break;                                     // uniform
        const int ebase = s.ncand;
        if (deg > 0) {
          int4 *rp = &x.rec[dd];
          *reinterpret_cast<int2 *>(rp) = make_int2(ebase + off, deg);
          rp->z = 0;                                          // arcs admitted by this expansion
        }
        __syncthreads();
        for (int j = tid; j < total; j += T) {
          int lo = 0, hi = T;
          while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (s.chunk_off[mid] <= j) lo = mid; else hi = mid;
          }
          const int a = s.chunk_ebeg[lo] + (j - s.chunk_off[lo]);
          const int owner = s.chunk_d[lo];
          const int4 arc = __ldg(&g.ne_arcs[a]);
          const float tot = s.chunk_cost[lo] + __int_as_float(arc.y);
          int4 entry = make_int4(0, kInfBits, a, -1);
          if (tot < cutoff) {
            bool created; int idx = 0;
            int ds = hash_insert_x(ctx, arc.x, &created, &idx);
            if (ds >= 0) {
              if (created) {
                hash[ds].z = idx;                             // provisional dense index (>= N1)
                if (idx < p.max_tpf) x.rec[idx] = make_int4(0, 0, 0, kInfBits);
              }
              entry.x = __float_as_int(tot); entry.y = arc.y; entry.w = ds;
              atomicAdd(&x.rec[owner].z, 1);
              uint32_t nv = f2ord(tot);
              uint32_t old = atomicMin(reinterpret_cast<uint32_t *>(&hash[ds].y), nv);
              if (nv < old && arc.w >= 0) {
                if (atomicExch(&hash[ds].w, stamp) != stamp) {
                  int q = atomicAdd(&s.wl_n[cur ^ 1], 1);
                  if (q < p.max_tpf) out[q] = ds;
                  else B2K_SET_ERR(s, B2K_ERR_OVERFLOW);
                }
              }
            }
          }
          x.adj[ebase + j] = entry;
          x.adjo[ebase + j] = owner;
        }
        __syncthreads();
      }
      __syncthreads();
      if (tid == 0) {
        s.wl_n[cur] = 0;
        s.stamp = stamp;
        if (s.wl_n[cur ^ 1] > p.max_tpf) s.wl_n[cur ^ 1] = p.max_tpf;
        s.cont = (s.wl_n[cur ^ 1] > 0 && !s.err);
      }
      cur ^= 1;
      first_round = false;
      __syncthreads();
    }
  }
  __syncthreads();
  B2K_TICK(s, 5);
  const int Nall = min(s.ntok_new, p.max_tpf);
  const int E = s.err ? 0 : min(s.q_n, p.adj_cap);
  const float kInfF = __int_as_float(0x7f800000);
  // dest hash slot -> dest dense index (superseded records are converted too; harmless)
  // (.x held source cost + weight.)  An arc can only ever fire in the replay if
  // final_cost(src) + w < cost(dest) after ProcessEmitting: the source is never cheaper
  // than its final cost and the destination never dearer than its initial one.  Arcs
  // that cannot fire are disabled for the walk (weight := +inf; the links pass does
  // not read the weight).
  for (int e = tid; e < E; e += T) {
    int4 en = x.adj[e];
    if (en.w < 0) continue;
    const int jd = hash[en.w].z;
    const float c0 = __int_as_float(x.rec[jd].w);
    if (!(__int_as_float(en.x) < c0)) x.adj[e].y = 0x7f800000;
    x.adj[e].x = jd;
  }
  __syncthreads();
  if (tid == 0) cs->arcs_ne += (unsigned long long)E;        // eps arcs examined by the closure
  // The replay only has to reproduce the order in which tokens are CREATED.  A token
  // from which no eps-created token is reachable through admitted arcs can never
  // influence that order (its pops create nothing, and it only ever modifies its own
  // descendants), so the walk is restricted to the ancestors of the created tokens:
  // backward reachability over the final records, then arcs into the rest are disabled.
  int *mark = wl1;                                           // the closure's worklists are idle now
  int qcarry = 0;
  int rescatter_from = N1;                                   // bucket arrays already hold tokens [0, N1)
  extern __shared__ __align__(16) unsigned char dyn_smem_base[];
  if (Nall > N1 && !s.err) {
    for (int d = tid; d < Nall; d += T) { mark[d] = (d >= N1); if (d >= N1) x.newseq[d - N1] = -1; }
    __syncthreads();
    for (int it = 0; it < 1000000; it++) {
      if (tid == 0) s.cont = 0;
      __syncthreads();
      for (int e = tid; e < E; e += T) {
        const int4 en = x.adj[e];
        if (en.w < 0 || en.y == 0x7f800000 || !mark[en.x]) continue;
        const int o = x.adjo[e];
        if (mark[o]) continue;
        const int4 r = x.rec[o];
        if (e < r.x || e >= r.x + r.y) continue;             // superseded record
        mark[o] = 1;
        s.cont = 1;
      }
      __syncthreads();
      const int again = s.cont;
      __syncthreads();
      if (!again) break;
    }
    for (int e = tid; e < E; e += T) {
      const int4 en = x.adj[e];
      if (en.w >= 0 && !mark[en.x]) x.adj[e].y = 0x7f800000;   // the links pass needs only arc id and dest slot
    }
    // initial worklist (:852-856) = the emitting tokens in list order, restricted to the
    // marked tokens whose final record admits something: collect them with their list
    // keys (first insertion index of the bucket, own insertion index) and sort the keys
    // in shared memory; only when they do not fit is the full list order built.
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(dyn_smem_base);
    const int key_cap = 1 << (31 - __clz(max(1, min(4096, p.rs_rcap + p.rs_ecap))));   // 8 bytes each, inside the walk's (not yet filled) arrays; a power of two
    if (tid == 0) s.rs_n = 0;
    __syncthreads();
    for (int d = tid; d < N1; d += T) {
      if (!mark[d] || x.rec[d].z <= 0) continue;
      int q = atomicAdd(&s.rs_n, 1);
      if (q < key_cap) keys[q] = ((unsigned long long)(uint32_t)x.bk[x.xb[d]].x << 32) | (uint32_t)d;
    }
    __syncthreads();
    qcarry = s.rs_n;
    if (qcarry <= key_cap) {
      int P = 1;
      while (P < qcarry) P <<= 1;
      for (int i = qcarry + tid; i < P; i += T) keys[i] = ~0ull;
      __syncthreads();
      for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = tid; i < P; i += T) {
            int ixj = i ^ j;
            if (ixj > i) {
              unsigned long long a = keys[i], b = keys[ixj];
              if ((a > b) == ((i & k) == 0)) { keys[i] = b; keys[ixj] = a; }
            }
          }
          __syncthreads();
        }
      for (int i = tid; i < qcarry; i += T) x.queue[i] = (int)(uint32_t)keys[i];
      __syncthreads();
    } else {
      order_finish<T>(N1, x, s);                             // (clears the bucket arrays: redone below)
      rescatter_from = 0;
      qcarry = 0;
      for (int base = 0; base < N1; base += T) {
        int k = base + tid, flag = 0, d = 0;
        if (k < N1) {
          d = x.order[k];
          flag = mark[d] && x.rec[d].z > 0;
        }
        int total;
        int excl = block_excl_scan<T>(flag, s.redi, &total);
        if (flag) x.queue[qcarry + excl] = d;
        qcarry += total;
      }
    }
  }
  __syncthreads();
  B2K_TICK(s, 6);
  // ---- the walk touches few tokens (the marked sources and their destinations) and
  // few arcs (the enabled ones): renumber the tokens compactly, compact each record
  // to its enabled arcs (order preserved: one warp per record), and run the LIFO walk
  // out of shared memory, so that each of its dependent steps costs a shared-memory
  // access instead of an L2/HBM round trip.  Falls back to the walk over the global
  // records when the set does not fit.
  unsigned char *dyn_smem = dyn_smem_base;
  // per token {replay cost, record offset | count << 16}; per arc {weight, dest id}: one 8-byte load each
  float2 *tk_s = reinterpret_cast<float2 *>(dyn_smem);
  float2 *en_s = tk_s + p.rs_rcap;
  unsigned short *ns_s = reinterpret_cast<unsigned short *>(en_s + p.rs_ecap);
  unsigned short *q_s = ns_s + p.rs_rcap;
  int *cid = x.sbase;                                        // idle between the two order_tokens calls
  int *dof = wl0;                                            // compact id -> dense index (worklists are idle now)
  bool replay_done = false;
  if (p.rs_rcap > 0 && !s.err && Nall > N1 && qcarry <= p.rs_qcap) {
    if (tid == 0) { s.rs_n = 0; s.rs_e = 0; s.rs_ok = 1; }
    for (int d = tid; d < Nall; d += T) cid[d] = -1;
    __syncthreads();
    auto claim = [&](int d) {
      if (atomicCAS(&cid[d], -1, -2) != -1) return;
      int id = atomicAdd(&s.rs_n, 1);
      if (id < p.rs_rcap) {
        tk_s[id] = make_float2(__int_as_float(x.rec[d].w), __int_as_float(0));
        ns_s[id] = 0xffff;
        dof[id] = d;
        cid[d] = id;
      } else {
        s.rs_ok = 0;
      }
    };
    for (int k = tid; k < qcarry; k += T) claim(x.queue[k]);
    for (int e = tid; e < E; e += T) {
      int4 en = x.adj[e];
      if (en.w < 0 || en.y == 0x7f800000) continue;
      int o = x.adjo[e];
      if (!mark[o]) continue;
      if (cid[en.x] == -1) claim(en.x);
      if (cid[o] == -1) claim(o);
    }
    __syncthreads();
    if (s.rs_ok) {
      const int R = s.rs_n;
      const int lane_id = tid & 31;
      // small records (the vast majority): one thread each; hub records: one warp each
      if (tid == 0) s.ncand = 0;
      __syncthreads();
      int *big_list = x.run;                                 // idle between the two list-order passes
      for (int id = tid; id < R; id += T) {
        const int d = dof[id];
        const int4 r = x.rec[d];
        if (!mark[d] || r.z == 0) continue;                  // destination only
        if (r.y > 8) { big_list[atomicAdd(&s.ncand, 1)] = id; continue; }
        int4 en[8];
        int total = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          en[i] = (i < r.y) ? x.adj[r.x + i] : make_int4(0, 0x7f800000, 0, -1);
          total += (en[i].w >= 0 && en[i].y != 0x7f800000);
        }
        if (!total) continue;
        const int eo = atomicAdd(&s.rs_e, total);
        if (eo + total > p.rs_ecap) { s.rs_ok = 0; continue; }
        int pos = eo;
#pragma unroll
        for (int i = 0; i < 8; i++)
          if (en[i].w >= 0 && en[i].y != 0x7f800000)
            en_s[pos++] = make_float2(__int_as_float(en[i].y), __int_as_float(cid[en[i].x]));
        tk_s[id].y = __int_as_float(eo | (total << 16));
      }
      __syncthreads();
      const int nbig = s.ncand;
      for (int bi = tid >> 5; bi < nbig; bi += T / 32) {     // one warp per hub record
        const int id = big_list[bi];
        const int d = dof[id];
        const int4 r = x.rec[d];
        int total = 0;
        for (int b0 = 0; b0 < r.y; b0 += 32) {
          int i = b0 + lane_id;
          int4 en = (i < r.y) ? x.adj[r.x + i] : make_int4(0, 0x7f800000, 0, -1);
          total += __popc(__ballot_sync(0xffffffffu, en.w >= 0 && en.y != 0x7f800000));
        }
        if (!total) continue;
        int eo = 0;
        if (lane_id == 0) eo = atomicAdd(&s.rs_e, total);
        eo = __shfl_sync(0xffffffffu, eo, 0);
        if (eo + total > p.rs_ecap || total > 0xffff) { if (lane_id == 0) s.rs_ok = 0; continue; }
        int pos = eo;
        for (int b0 = 0; b0 < r.y; b0 += 32) {
          int i = b0 + lane_id;
          int4 en = (i < r.y) ? x.adj[r.x + i] : make_int4(0, 0x7f800000, 0, -1);
          bool keep = en.w >= 0 && en.y != 0x7f800000;
          uint32_t m = __ballot_sync(0xffffffffu, keep);
          if (keep) {
            int q = pos + __popc(m & ((1u << lane_id) - 1u));
            en_s[q] = make_float2(__int_as_float(en.y), __int_as_float(cid[en.x]));
          }
          pos += __popc(m);
        }
        if (lane_id == 0) tk_s[id].y = __int_as_float(eo | (total << 16));
      }
      for (int k = tid; k < qcarry; k += T) q_s[k] = (unsigned short)cid[x.queue[k]];
    }
    __syncthreads();
    if (s.rs_ok && tid == 0) {
      int qn = qcarry, next = 0;
      bool ok = true;
      int npop = 0, nvis = 0;
      while (qn > 0) {
        const int d = q_s[--qn];
        const float2 td = tk_s[d];
        const float c = td.x;
        npop++;
        if (c >= cutoff) continue;
        const int oc = __float_as_int(td.y);
        const int e0 = oc & 0xffff, e1 = e0 + (int)((unsigned)oc >> 16);
        nvis += e1 - e0;
        for (int e = e0; e < e1; e++) {
          const float2 ent = en_s[e];
          const float tot = c + ent.x;
          if (tot < cutoff) {
            const int j = __float_as_int(ent.y);
            const float2 tj = tk_s[j];
            if (tot < tj.x) {
              tk_s[j].x = tot;
              if (tj.x == kInfF) ns_s[j] = (unsigned short)next++;
              if ((unsigned)__float_as_int(tj.y) >> 16) {
                if (qn < p.rs_qcap) q_s[qn++] = (unsigned short)j;
                else { ok = false; qn = 0; break; }
              }
            }
          }
        }
      }
      if (ok) {
        if (next != Nall - N1) B2K_SET_ERR(s, B2K_ERR_STATE);
        s.prof[12] += (unsigned long long)npop; s.prof[13] += (unsigned long long)nvis; s.prof[14] += 1;
      } else {
        s.rs_ok = 0;                                         // worklist outgrew shared memory: redo below
      }
    }
    __syncthreads();
    if (s.rs_ok) {
      replay_done = true;
      if (!s.err)
        for (int d = N1 + tid; d < Nall; d += T) {
          int id = cid[d];
          if (id >= 0) x.newseq[d - N1] = (int)ns_s[id];
          else B2K_SET_ERR(s, B2K_ERR_STATE);                // an eps-created token is always some record's destination
        }
    }
    __syncthreads();
  }
  // literal LIFO replay by one thread over the dense records in global memory
  if (tid == 0 && !s.err && !replay_done && Nall > N1) {
    int qn = qcarry, next = 0;
    int npop = 0, nvis = 0;
    while (qn > 0) {
      const int d = x.queue[--qn];
      const int4 r = x.rec[d];
      const float c = __int_as_float(r.w);
      npop++;
      if (c >= cutoff || r.z == 0) continue;
      nvis += r.y;
      for (int e = r.x; e < r.x + r.y; e++) {
        const int4 en = x.adj[e];
        const float tot = c + __int_as_float(en.y);
        if (tot < cutoff) {
          const int j = en.x;
          const int4 rj = x.rec[j];
          const float old = __int_as_float(rj.w);
          if (tot < old) {                                   // FindOrAddToken: new or improved -> changed
            x.rec[j].w = __float_as_int(tot);
            if (old == kInfF) x.newseq[j - N1] = next++;
            if (rj.z > 0) {
              if (qn < p.queue_cap) x.queue[qn++] = j;
              else { B2K_SET_ERR(s, B2K_ERR_OVERFLOW); break; }
            }
          }
        }
      }
    }
    if (!s.err && next != Nall - N1) B2K_SET_ERR(s, B2K_ERR_STATE);
    s.prof[12] += (unsigned long long)npop; s.prof[13] += (unsigned long long)nvis;
  }
  __syncthreads();
  B2K_TICK(s, 7);
  // final insertion index of the eps-created tokens
  if (!s.err) {
    for (int d = N1 + tid; d < Nall; d += T) {
      int slot = ctx.tokslot[d];
      int ins = N1 + x.newseq[d - N1];
      hash[slot].z = ins;
      x.by_ins[ins] = slot;
    }
  }
  __syncthreads();
  B2K_TICK(s, 8);
  const int N = Nall;
  bucket_scatter<T>(rescatter_from, N, Hc, hash, x);
  int *order_slot = x.queue;                                 // (the replay worklist is done) list rank -> hash slot
  order_finish<T>(N, x, s, order_slot);
  B2K_TICK(s, 9);
  // eps links = the admitted entries of the final records
  if (!s.err) {
    const int lane_id = tid & 31;
    for (int base = 0; base < E; base += T) {
      const int e = base + tid;
      bool live = false;
      int4 en = make_int4(0, 0, 0, -1);
      int src = 0;
      if (e < E) {
        en = x.adj[e];
        if (en.w >= 0) {
          const int o = x.adjo[e];
          const int4 r = x.rec[o];
          live = (e >= r.x && e < r.x + r.y);
          src = (o < N1) ? x.by_ins[o] : ctx.tokslot[o];
        }
      }
      const uint32_t m = __ballot_sync(0xffffffffu, live);
      if (m) {
        int lb = 0;
        if (lane_id == 0) lb = atomicAdd(&s.nlink_new, __popc(m));
        lb = __shfl_sync(0xffffffffu, lb, 0);
        if (live) {
          int li = lb + __popc(m & ((1u << lane_id) - 1u));
          if (lbase + li < p.max_links)
            links[lbase + li] = make_int4(src, en.w, (int)((uint32_t)en.z | B2K_EPS_FLAG), 0);
          else
            B2K_SET_ERR(s, B2K_ERR_OVERFLOW);
        }
      }
    }
  }
  __syncthreads();
  B2K_TICK(s, 10);
  const int nlink = s.nlink_new;
  const bool fits = (ctx.tbase + N <= p.max_tokens);
  if (!fits && tid == 0) B2K_SET_ERR(s, B2K_ERR_OVERFLOW);
  __syncthreads();
  if (!s.err) {
    for (int r = tid; r < N; r += T) {
      const int slot = order_slot[r];
      const int4 hs = hash[slot];
      tok_state[ctx.tbase + r] = hs.x;
      tok_cost[ctx.tbase + r] = ord2f((uint32_t)hs.y);
      hash[slot].w = r;
    }
  }
  __syncthreads();
  if (!s.err) {
    for (int li = tid; li < nlink; li += T) {
      int4 lk = links[lbase + li];
      if ((uint32_t)lk.z & B2K_EPS_FLAG) lk.x = ctx.tbase + hash[lk.x].w;
      lk.y = ctx.tbase + hash[lk.y].w;
      links[lbase + li] = lk;
    }
  }
  __syncthreads();
  for (int i = tid; i < N; i += T)
    hash[ctx.tokslot[i]] = make_int4(B2K_HASH_EMPTY, (int)B2K_INF_ORD, 0x7fffffff, 0);
  if (tid == 0) {
    size_t fo = (size_t)ch * (p.max_frames + 2);
    p.frame_tok_begin[fo + list_index] = ctx.tbase;
    p.frame_tok_begin[fo + list_index + 1] = ctx.tbase + N;
    p.frame_link_begin[fo + list_index] = lbase;
    p.frame_link_eps[fo + list_index] = lbase + n_emit_links;
    p.frame_link_begin[fo + list_index + 1] = lbase + nlink;
    if (list_index > 0) {
      size_t co = (size_t)ch * (p.max_frames + 1) + (list_index - 1);
      p.frame_cost_offset[co] = cost_offset;
      p.frame_cutoff[co] = cutoff;
    }
    B2K_TICK(s, 11);
  }
  __syncthreads();
}

template <int T, bool PROF>
__device__ void dec_advance_exact_lane(const DecParams &p, DecShared<T> &s, const int lane, const int slot) {
  constexpr int IT = 4;
  const int tid = threadIdx.x;
  const int ch = p.lane_channel[lane];
  ChanState *cs = &p.chan[ch];
  const FstDev &g = p.fst;
  const float kInf = __int_as_float(0x7f800000);

  if (cs->status != B2K_OK) return;
  if (!p.do_init && cs->frames_decoded < 0) {
    if (tid == 0) { cs->status = B2K_ERR_STATE; cs->err_line = __LINE__; }
    return;
  }
  int32_t *tok_state = p.tok_state + (size_t)ch * p.max_tokens;
  float *tok_cost = p.tok_cost + (size_t)ch * p.max_tokens;
  int4 *links = p.links + (size_t)ch * p.max_links;

  LaneCtx ctx;
  ctx.hash = p.hash + (size_t)slot * p.hash_size;
  ctx.tokslot = p.tokslot + (size_t)slot * p.max_tpf;
  ctx.tok_state = tok_state;
  ctx.hash_mask = p.hash_size - 1;
  ctx.hash_log = p.hash_log;
  ctx.max_tpf = p.max_tpf;
  ctx.max_tokens = p.max_tokens;
  ctx.ntok_new = &s.ntok_new;
  ctx.err = &s.err;
  ctx.err_line = &s.err_line;
  set_l1(ctx, 0);
  XScratch x;
  x.bm = p.x_bm + (size_t)slot * (p.pos_cap / 32);
  x.wbase = p.x_wbase + (size_t)slot * (p.pos_cap / 32);
  x.by_ins = p.x_by_ins + (size_t)slot * p.max_tpf;
  x.bk = p.x_bk + (size_t)slot * p.hc_cap;
  x.sbase = p.x_sbase + (size_t)slot * p.max_tpf;
  x.run = p.x_run + (size_t)slot * p.max_tpf;
  x.order = p.x_order + (size_t)slot * p.max_tpf;
  x.queue = p.cand + (size_t)slot * 5 * p.cand_cap;     // idle in this mode
  x.xb = p.x_xb + (size_t)slot * p.max_tpf;
  x.rec = p.x_rec + (size_t)slot * p.max_tpf;
  x.newseq = p.x_newseq + (size_t)slot * p.max_tpf;
  x.adj = p.x_adj + (size_t)slot * p.adj_cap;
  x.adjo = p.x_adjo + (size_t)slot * p.adj_cap;

  if (tid == 0) { s.err = 0; s.err_line = 0; s.stamp = 0; }
  __syncthreads();

  if (p.do_init) {
    if (tid == 0) { s.ntok_new = 0; s.nlink_new = 0; }
    __syncthreads();
    ctx.tbase = 0;
    if (tid == 0) {
      bool created; int idx = 0;
      int slot = hash_insert_x(ctx, g.start, &created, &idx);
      if (slot >= 0) { ctx.hash[slot].y = (int)f2ord(0.0f); ctx.hash[slot].z = 0; x.by_ins[0] = slot; }
    }
    finish_frame_exact<T, PROF>(p, s, ctx, x, slot, ch, 0, p.beam, 0.0f, 0, 1000, cs);
    if (tid == 0) {
      cs->frames_decoded = 0;
      cs->ntok = min(s.ntok_new, p.max_tpf);
      cs->nlink = s.nlink_new;
      cs->finalized = 0;
      cs->hc = 1000;                                   // toks_.SetSize(1000) (:39)
      if (s.err) { cs->status = s.err; cs->err_line = s.err_line; }
    }
    if (s.err) reset_lane_hash<T>(ctx.hash, p.hash_size);
    return;
  }

  const int nframes = p.lane_nframes[lane];
  const float *ll_base = p.lane_loglikes[lane];
  int frames_decoded = cs->frames_decoded;
  int32_t tbase = cs->ntok, lbase = cs->nlink;
  int Hc = cs->hc;
  unsigned long long arcs_e_total = 0;
  const size_t fo = (size_t)ch * (p.max_frames + 2);

  if (tid == 0) { for (int i = 0; i < 16; i++) s.prof[i] = 0; s.tlast = PROF ? clock64() : 0; }
  const long long t_kernel0 = PROF ? clock64() : 0;
  extern __shared__ __align__(16) unsigned char dyn_smem_base[];
  float *ll_s = reinterpret_cast<float *>(dyn_smem_base + p.rs_bytes);
  for (int fi = 0; fi < nframes; fi++) {
    if (frames_decoded >= p.max_frames) { if (tid == 0) B2K_SET_ERR(s, B2K_ERR_OVERFLOW); __syncthreads(); break; }
    const float *llg = ll_base + (size_t)fi * p.row_stride;
    // the frame's log-likelihood row is gathered once per arc: stage it in shared memory (the block
    // reductions of the cutoff below are the barrier between this fill and the first reader)
    if (p.ll_smem) for (int i = tid; i < p.num_pdfs; i += T) ll_s[i] = __ldcs(&llg[i]);
    const int pb = p.frame_tok_begin[fo + frames_decoded];
    const int pe = p.frame_tok_begin[fo + frames_decoded + 1];
    const int K = pe - pb;
    set_l1(ctx, K);

    // ---- GetCutoff (:653-720); ties -> first in list order (strict < at :662/:675)
    unsigned long long local = ~0ull;
    for (int i = pb + tid; i < pe; i += T) {
      unsigned long long key = ((unsigned long long)f2ord(tok_cost[i]) << 32) | (uint32_t)(i - pb);
      local = key < local ? key : local;
    }
    unsigned long long bestkey = block_min_u64<T>(local, s.red64);
    float best_cost = kInf;
    int best_state = -1;
    if (K > 0) {
      best_cost = ord2f((uint32_t)(bestkey >> 32));
      best_state = tok_state[pb + (int)(uint32_t)(bestkey & 0xffffffffu)];
    }
    const float beam_cutoff = best_cost + p.beam;
    float cur_cutoff = beam_cutoff, adaptive_beam = p.beam;
    if (K > 0 && !(p.max_active == 0x7fffffff && p.min_active == 0)) {
      int c_lt = 0, c_le = 0;
      for (int i = pb + tid; i < pe; i += T) {
        float c = tok_cost[i];
        c_lt += (c < beam_cutoff);
        c_le += (c <= beam_cutoff);
      }
      c_lt = block_sum_i32<T>(c_lt, s.redi);
      c_le = block_sum_i32<T>(c_le, s.redi);
      bool done = false;
      if (K > p.max_active && c_lt > p.max_active) {
        float mac = block_select_kth<T>(tok_cost + pb, K, p.max_active, s.hist, s.pref);
        cur_cutoff = mac;
        adaptive_beam = mac - best_cost + p.beam_delta;
        done = true;
      }
      if (!done) {
        float min_active_cutoff = kInf;
        if (K > p.min_active) {
          if (p.min_active == 0) min_active_cutoff = best_cost;
          else if (c_le <= p.min_active)
            min_active_cutoff = block_select_kth<T>(tok_cost + pb, K, p.min_active, s.hist, s.pref);
          else min_active_cutoff = beam_cutoff;
        }
        if (min_active_cutoff > beam_cutoff) {
          adaptive_beam = min_active_cutoff - best_cost + p.beam_delta;
          cur_cutoff = min_active_cutoff;
        }
      }
    }
    // PossiblyResizeHash(tok_cnt) (:227-233)
    {
      long long new_sz = (long long)((float)K * p.hash_ratio);
      if (new_sz > (long long)Hc) Hc = (int)min(new_sz, (long long)0x7fffffff);
      if (Hc > p.hc_cap) { if (tid == 0) B2K_SET_ERR(s, B2K_ERR_OVERFLOW); __syncthreads(); break; }
    }
    const float cost_offset = (K > 0) ? -best_cost : 0.0f;

    // ---- seed (:753-768)
    uint32_t seed_local = B2K_INF_ORD;
    if (K > 0) {
      int2 o0 = __ldg(&g.st_off[best_state]), o1 = __ldg(&g.st_off[best_state + 1]);
      for (int a = o0.x + tid; a < o1.x; a += T) {
        int4 arc = __ldg(&g.e_arcs[a]);
        float new_weight = __int_as_float(arc.y) + cost_offset - (p.ll_smem ? ll_s[arc.z] : __ldg(&llg[arc.z])) + best_cost;
        seed_local = min(seed_local, f2ord(new_weight + adaptive_beam));
      }
    }
    const float seed_cutoff = ord2f(block_min_u32<T>(seed_local, s.red32));

    // ---- main loop (:779-812): admission against the exclusive prefix-min
    if (tid == 0) { s.ntok_new = 0; s.nlink_new = 0; }
    __syncthreads();
    B2K_TICK(s, 0);
    ctx.tbase = tbase;
    float carry = kInf;
    int pos_base = 0;
    int round_no = 0;
    // Tokens are expanded in chunks of CT = T * TPT: the chunk's out-degrees are scanned
    // once, then its arcs are processed ARC-parallel in list order (rounds of T * IT arcs).
    constexpr int TPT = DecShared<T>::TPT, CT = DecShared<T>::CT;
    for (int cb = pb; cb < pe; cb += CT) {
#pragma unroll
      for (int u = 0; u < TPT; u++) {                       // striped (coalesced) loads
        const int q = u * T + tid, i = cb + q;
        int deg = 0, ebeg = 0;
        float c = 0.f;
        if (i < pe) {
          c = tok_cost[i];
          if (c <= cur_cutoff) {
            int st = tok_state[i];
            int2 o0 = __ldg(&g.st_off[st]), o1 = __ldg(&g.st_off[st + 1]);
            ebeg = o0.x;
            deg = o1.x - o0.x;
          }
        }
        s.chunk_off[q] = deg; s.chunk_ebeg[q] = ebeg; s.chunk_cost[q] = c;
      }
      __syncthreads();
      int total;
      {                                                      // blocked exclusive scan of the degrees
        int dloc[TPT], sum = 0;
#pragma unroll
        for (int u = 0; u < TPT; u++) { dloc[u] = s.chunk_off[tid * TPT + u]; sum += dloc[u]; }
        int off = block_excl_scan<T>(sum, s.redi, &total);
#pragma unroll
        for (int u = 0; u < TPT; u++) { s.chunk_off[tid * TPT + u] = off; off += dloc[u]; }
      }
      if (tid == 0) { s.chunk_off[CT] = total; arcs_e_total += (unsigned long long)total; }
      __syncthreads();
      for (int r0 = 0; r0 < total; r0 += T * IT, round_no++) {
        const int j0 = r0 + tid * IT;
        float tots[IT], acs[IT];
        int arcid[IT], nexts[IT], srcs[IT];
        int lo = 0;
        if (j0 < total) {
          int hi = CT;
          while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (s.chunk_off[mid] <= j0) lo = mid; else hi = mid;
          }
        }
#pragma unroll
        for (int k = 0; k < IT; k++) {
          int j = j0 + k;
          tots[k] = kInf; acs[k] = 0.f; arcid[k] = 0; nexts[k] = 0; srcs[k] = 0;
          if (j < total) {
            while (s.chunk_off[lo + 1] <= j) lo++;
            int a = s.chunk_ebeg[lo] + (j - s.chunk_off[lo]);
            int4 arc = __ldg(&g.e_arcs[a]);
            float ac = cost_offset - (p.ll_smem ? ll_s[arc.z] : __ldg(&llg[arc.z]));
            tots[k] = s.chunk_cost[lo] + ac + __int_as_float(arc.y);
            acs[k] = ac; arcid[k] = a; nexts[k] = arc.x; srcs[k] = cb + lo;
          }
        }
        float ex[IT];
        float pm = kInf;
#pragma unroll
        for (int k = 0; k < IT; k++) { ex[k] = pm; pm = fminf(pm, tots[k]); }
        // block-wide exclusive scan (min) of pm in thread order
        const int lane_id = tid & 31, warp = tid >> 5;
        float incl = pm;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          float n = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane_id >= o) incl = fminf(incl, n);
        }
        float wexcl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane_id == 0) wexcl = kInf;
        float *buf = s.scanf_[round_no & 1];
        if (lane_id == 31) buf[warp] = incl;
        __syncthreads();
        float wpre = kInf, tot_all = kInf;
#pragma unroll
        for (int w = 0; w < T / 32; w++) {
          float v = buf[w];
          if (w < warp) wpre = fminf(wpre, v);
          tot_all = fminf(tot_all, v);
        }
        const float base = fminf(carry, fminf(wpre, wexcl));
#pragma unroll
        for (int k = 0; k < IT; k++) {
          int j = j0 + k;
          float excl = fminf(base, ex[k]);
          float rc = fminf(seed_cutoff, excl + adaptive_beam);
          bool adm = (j < total) && (tots[k] < rc);
          int slot = -1;
          if (adm) {
            bool created; int idx;
            slot = hash_insert_x(ctx, nexts[k], &created, &idx);
            if (slot >= 0) {
              atomicMin(reinterpret_cast<uint32_t *>(&ctx.hash[slot].y), f2ord(tots[k]));
              atomicMin(&ctx.hash[slot].z, pos_base + j);
            } else adm = false;
          }
          uint32_t m = __ballot_sync(0xffffffffu, adm);
          if (m) {
            int lb = 0;
            if (lane_id == 0) lb = atomicAdd(&s.nlink_new, __popc(m));
            lb = __shfl_sync(0xffffffffu, lb, 0);
            if (adm) {
              int li = lb + __popc(m & ((1u << lane_id) - 1u));
              if (lbase + li < p.max_links)
                links[lbase + li] = make_int4(srcs[k], slot, arcid[k], __float_as_int(acs[k]));
              else
                B2K_SET_ERR(s, B2K_ERR_OVERFLOW);
            }
          }
        }
        carry = fminf(carry, tot_all);
      }
      pos_base += total;
      __syncthreads();
    }
    const float next_cutoff = fminf(seed_cutoff, carry + adaptive_beam);
    __syncthreads();
    B2K_TICK(s, 1);
    // ---- insertion index of the tokens created above = rank of their first
    //      admitted position (bitmap rank)
    const int N1 = min(s.ntok_new, p.max_tpf);
    if (pos_base > p.pos_cap) { if (tid == 0) B2K_SET_ERR(s, B2K_ERR_OVERFLOW); }
    __syncthreads();
    if (!s.err) {
      const int W = (pos_base + 31) >> 5;
      for (int w = tid; w < W; w += T) x.bm[w] = 0u;
      __syncthreads();
      for (int i = tid; i < N1; i += T) {
        int seq = ctx.hash[ctx.tokslot[i]].z;
        x.sbase[i] = seq;                                     // (scratch: idle until order_finish)
        atomicOr(&x.bm[seq >> 5], 1u << (seq & 31));
      }
      __syncthreads();
      int wc = 0;
      for (int base = 0; base < W; base += T) {
        int w = base + tid;
        int cnt = (w < W) ? __popc(x.bm[w]) : 0;
        int total;
        int excl = block_excl_scan<T>(cnt, s.redi, &total);
        if (w < W) x.wbase[w] = wc + excl;
        wc += total;
      }
      __syncthreads();
      for (int i = tid; i < N1; i += T) {
        int slot = ctx.tokslot[i];
        int seq = x.sbase[i];
        int ins = x.wbase[seq >> 5] + __popc(x.bm[seq >> 5] & ((1u << (seq & 31)) - 1u));
        ctx.hash[slot].z = ins;
        x.by_ins[ins] = slot;
      }
    }
    B2K_TICK(s, 2);
    finish_frame_exact<T, PROF>(p, s, ctx, x, slot, ch, frames_decoded + 1, next_cutoff, cost_offset, lbase,
                          Hc, cs);
    if (s.err) break;
    tbase += min(s.ntok_new, p.max_tpf);
    lbase += s.nlink_new;
    frames_decoded++;
    __syncthreads();
  }
  if (tid == 0) {
    cs->frames_decoded = frames_decoded;
    cs->ntok = tbase;
    cs->nlink = lbase;
    cs->arcs_e += arcs_e_total;
    cs->hc = Hc;
    if (s.err) { cs->status = s.err; cs->err_line = s.err_line; }
    for (int i = 0; i < 15; i++) cs->prof[i] += s.prof[i];
    if (PROF) cs->prof[15] += (unsigned long long)(clock64() - t_kernel0);
  }
  if (s.err) {
    reset_lane_hash<T>(ctx.hash, p.hash_size);
    for (int i = tid; i < p.hc_cap; i += T) x.bk[i] = make_int4(0x7fffffff, 0, 0, 0);
  }
}

// PROF = per-phase cycle counters (bench.py's decoder_phase_share pass; B2K_DEC_PROF=1), compiled out otherwise
template <int T, bool PROF>
__global__ void __launch_bounds__(T, (T == 512 ? 2 : 1)) dec_advance_exact_kernel(DecParams p) {
  __shared__ DecShared<T> s;
  B2K_PERSISTENT_LANES((dec_advance_exact_lane<T, PROF>(p, s, lane_, (int)blockIdx.x)))
}

// ====================================================================== reference-order mode, second generation
//
// Same results as the frame step above (bit for bit: tests/test_decoder_gpu.py runs both), built around what the
// round-2 profile showed (profiles/r02_decoder_history.md): the first generation visits every token's hash slot ~9
// times and its HashList bucket record ~5 times per frame, reads the state table twice per token, and spends 46 % of
// its time in the epsilon phases although they examine 10 % of the arcs.  Here
//   * the token table is keyed by the reference's OWN bucket (state % hash size, hash-list-inl.h:130): all tokens of a
//     bucket lie on one probe sequence before its first empty slot, so the HashList order (first insertion of the
//     bucket, insertion within the bucket) is read off the table by walking that run -- no bucket array, no scatter,
//     no per-bucket atomics;
//   * a token's insertion key is the position of its first admitted arc (tokens created by ProcessNonemitting:
//     positions after all arcs, in replay order); list positions come from ONE exclusive scan over that key space;
//   * the thread that creates a token reads the state table once and leaves {emitting arcs, epsilon arcs} in a dense
//     per-frame record: the closure finds its epsilon sources there, and the commit hands the emitting ranges to the
//     next frame's expansion in list order, so the expansion never touches the state table;
//   * the cost a token had before the closure is reconstructed from the values atomicMin returned (the largest one a
//     successful lowering saw), so no per-token snapshot pass exists; a destination the closure never lowered cannot
//     fire in the replay at all;
//   * epsilon links are written with final arena indices; emitting links are remapped once.
// A slot is {state, cost (ordered bits), insertion key, creation index}.

struct X2 {
  int4 *trec;          // [max_tpf] creation index -> {emitting arc begin, emitting degree, eps arc begin, eps degree}
  int4 *rec;           // [max_tpf] closure record {offset, eps degree, #admitted, replay cost bits (global walk only)}
  uint32_t *c0;        // [max_tpf] cost before the closure (ordered bits) if the closure lowered it; 0 = never lowered; +inf = created by the closure
  int32_t *newseq;     // [max_tpf]
  int4 *adj; int32_t *adjo; uint32_t *adjc;   // [adj_cap] eps arc entries, their owners, the cost the destination had before this entry lowered it (0 = did not)
  int32_t *wl0, *wl1;  // [max_tpf]
  int32_t *queue;      // [5 * cand_cap]
  uint32_t *qstamp;    // [hash_size] re-queue filter of the closure rounds
  int4 *tok4;          // [max_tpf] creation index -> {state, cost bits, first key of the bucket, rank inside the bucket}
  uint32_t *hw;        // [pos_cap + max_tpf + 32] key -> bucket population at bucket heads, then its exclusive scan
  int32_t *rank;       // [max_tpf] creation index -> list rank
  int32_t *pf_ebeg, *pf_edeg;   // [max_tpf] list rank -> emitting arcs (input of the next frame's expansion)
  int32_t *cid, *big;  // [max_tpf]
};

#define B2K_V2_L1_PROBES 32

// i-th slot of bucket b's probe sequence (see the note at probe_slot: positions are stable because slots are never
// freed inside a frame).  Level 1 = a window of the table sized from the reference's HashList size (>= 2x the previous
// token count), level 2 = the whole table.
__device__ __forceinline__ uint32_t probe_slot_b(const LaneCtx &c, uint32_t b, int i) {
  if (i < B2K_V2_L1_PROBES) return (((b * 2654435761u) >> (32 - c.l1_log)) + (uint32_t)i) & (uint32_t)c.l1_mask;
  return (((b * 0x85ebca6bu) >> (32 - c.hash_log)) + (uint32_t)(i - B2K_V2_L1_PROBES)) & (uint32_t)c.hash_mask;
}

// true if the level-2 part of bucket b's sequence passes over a slot of its own level-1 window (already visited)
__device__ __forceinline__ bool probe_revisits_b(const LaneCtx &c, uint32_t b, int i, uint32_t slot) {
  if (i < B2K_V2_L1_PROBES || slot > (uint32_t)c.l1_mask) return false;
  const uint32_t w0 = ((b * 2654435761u) >> (32 - c.l1_log)) & (uint32_t)c.l1_mask;
  return ((slot - w0) & (uint32_t)c.l1_mask) < (uint32_t)B2K_V2_L1_PROBES;
}

// state % Hc (hash-list-inl.h:130) without the generic 32-bit modulo (the source-level profile had 9 % of all stall samples
// on it): Lemire's fastmod, exact for every 32-bit numerator and divisor.  M = floor(2^64 / d) + 1 (0 for d = 1).
__device__ __forceinline__ uint32_t bucket_b(const LaneCtx &c, int32_t state) {
  const unsigned long long low = c.hc_m * (unsigned long long)(uint32_t)state;
  return (uint32_t)__umul64hi(low, (unsigned long long)c.hc);
}

__device__ __forceinline__ void set_l1_b(LaneCtx &c, uint32_t Hc, int shift) {
  c.hc = Hc;
  c.hc_m = 0xffffffffffffffffull / (unsigned long long)Hc + 1ull;
  int lg = 12;
  while ((1ull << lg) < ((unsigned long long)Hc << shift) && lg < c.hash_log) lg++;
  c.l1_log = lg;
  c.l1_mask = (1 << lg) - 1;
}

// find-or-insert on the bucket-keyed table; the creator records slot <-> creation index
__device__ __forceinline__ int hash_insert_b(const LaneCtx &c, int32_t state, uint32_t Hc, bool *created, int *idx_out) {
  const uint32_t b = bucket_b(c, state);
  *created = false;
  for (int probe = 0; probe <= c.hash_mask + B2K_V2_L1_PROBES; probe++) {
    const uint32_t h = probe_slot_b(c, b, probe);
    if (probe == B2K_V2_L1_PROBES) *c.used_l2 = 1;          // (the table scan of the list order needs every bucket inside one level-1 run)
    int *keyp = reinterpret_cast<int *>(&c.hash[h]);
    int old = atomicCAS(keyp, B2K_HASH_EMPTY, state);
    if (old == B2K_HASH_EMPTY) {
      int idx = atomicAdd(c.ntok_new, 1);
      if (idx < c.max_tpf) { c.tokslot[idx] = (int)h; keyp[3] = idx; }
      else do { if (atomicCAS(c.err, 0, B2K_ERR_OVERFLOW) == 0) *c.err_line = __LINE__; } while (0);
      *created = true;
      *idx_out = idx;
      return (int)h;
    }
    if (old == state) return (int)h;
  }
  do { if (atomicCAS(c.err, 0, B2K_ERR_OVERFLOW) == 0) *c.err_line = __LINE__; } while (0);
  return -1;
}

// in-place exclusive scan of a[0, n): every warp owns a contiguous range (coalesced), one block-level step
template <int T>
__device__ void block_excl_scan_array(uint32_t *a, int n, int *sh /*[T/32+1]*/) {
  constexpr int W = T / 32;
  const int warp = threadIdx.x >> 5, lane_id = threadIdx.x & 31;
  const int per = ((n + W * 32 - 1) / (W * 32)) * 32;
  const int b0 = min(n, warp * per), b1 = min(n, b0 + per);
  uint32_t sum = 0;
  for (int i = b0 + lane_id; i < b1; i += 32) sum += a[i];
  sum = __reduce_add_sync(0xffffffffu, sum);
  __syncthreads();
  if (lane_id == 0) sh[warp] = (int)sum;
  __syncthreads();
  uint32_t carry = 0;
  for (int w = 0; w < warp; w++) carry += (uint32_t)sh[w];
  for (int i0 = b0; i0 < b1; i0 += 32) {
    const int i = i0 + lane_id;
    const uint32_t v = (i < b1) ? a[i] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t nb = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane_id >= o) incl += nb;
    }
    if (i < b1) a[i] = carry + incl - v;
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  __syncthreads();
}

// ascending bitonic sort of keys[0, P2) (P2 a power of two; padding = ~0); keys may live in shared or global memory
template <int T>
__device__ void block_bitonic_sort_u64(unsigned long long *keys, int P2) {
  for (int k = 2; k <= P2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P2; i += T) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long a = keys[i], b = keys[ixj];
          if ((a > b) == ((i & k) == 0)) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
}

// ProcessNonemitting (closure, creation order by replay), list order, commit.  On entry the emitting pass has
// inserted N1 = s.ntok_new tokens (slot.z = first admitted position < P, slot.w = creation index, trec filled).
template <int T, bool PROF>
__device__ void finish_frame_v2(const DecParams &p, DecShared<T> &s, const LaneCtx &ctx, const X2 &x, int ch,
                                int list_index, float cutoff, float cost_offset, int32_t lbase, uint32_t Hc, int P,
                                ChanState *cs) {
  const int tid = threadIdx.x;
  const FstDev &g = p.fst;
  int4 *hash = ctx.hash;
  const int32_t *tokslot = ctx.tokslot;
  int4 *links = p.links + (size_t)ch * p.max_links;
  int32_t *tok_state = p.tok_state + (size_t)ch * p.max_tokens;
  float *tok_cost = p.tok_cost + (size_t)ch * p.max_tokens;
  const int kInfBits = 0x7f800000;
  const float kInfF = __int_as_float(kInfBits);
  __syncthreads();
  if (s.err) return;     // uniform: nothing writes err between the barrier and this read
  const int n_emit_links = s.nlink_new;
  const int N1 = min(s.ntok_new, p.max_tpf);
  B2K_TICK(s, 3);
  // ---- first worklist of the closure: the tokens that have eps arcs (dense records) and are below the cutoff
  int32_t *wl0 = x.wl0, *wl1 = x.wl1;
  int4 *wlx = reinterpret_cast<int4 *>(x.queue);              // idle until the replay worklist is built
  if (tid == 0) { s.wl_n[0] = 0; s.wl_n[1] = 0; s.q_n = 0; }
  __syncthreads();
  for (int base = 0; base < N1; base += T) {
    const int d = base + tid;
    int deg = 0, ebeg = 0, cbits = 0;
    if (d < N1) {
      const int4 t = x.trec[d];
      x.rec[d] = make_int4(0, 0, 0, 0);
      x.c0[d] = 0u;
      if (t.w > 0) {
        const float c = ord2f((uint32_t)hash[tokslot[d]].y);
        if (c < cutoff) { deg = t.w; ebeg = t.z; cbits = __float_as_int(c); }
      }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, deg > 0);
    if (m) {
      int wb = 0;
      if ((tid & 31) == 0) wb = atomicAdd(&s.wl_n[0], __popc(m));
      wb = __shfl_sync(0xffffffffu, wb, 0);
      if (deg > 0) wlx[wb + __popc(m & ((1u << (tid & 31)) - 1u))] = make_int4(cbits, ebeg, deg, d);
    }
  }
  __syncthreads();
  if (tid == 0) s.cont = (s.wl_n[0] > 0 && !s.err);
  __syncthreads();
  B2K_TICK(s, 4);
  // ---- closure by parallel relaxation, ARC-parallel per round (see the first generation for why the LAST record of a
  //      token is exactly its set of eps links and a superset of what the replay can admit)
  {
    int cur = 0;
    bool first_round = true;
    while (s.cont) {
      const int n = s.wl_n[cur];
      const int stamp = s.stamp + 1;
      int32_t *in = cur ? wl1 : wl0;
      int32_t *out = cur ? wl0 : wl1;
      for (int cb = 0; cb < n; cb += T) {
        const int k = cb + tid;
        int deg = 0, ebeg = 0, dd = 0;
        float c = 0.f;
        if (k < n) {
          if (first_round) {
            const int4 w = wlx[k];
            c = __int_as_float(w.x); ebeg = w.y; deg = w.z; dd = w.w;
          } else {
            const int4 hs = __ldcg(&hash[in[k]]);             // cost may be lowered concurrently: read at L2
            c = ord2f((uint32_t)hs.y);
            if (c < cutoff) {
              dd = hs.w;
              const int4 t = x.trec[dd];
              ebeg = t.z; deg = t.w;
            }
          }
        }
        int total;
        const int off = block_excl_scan<T>(deg, s.redi, &total);
        s.chunk_off[tid] = off; s.chunk_ebeg[tid] = ebeg; s.chunk_cost[tid] = c; s.chunk_d[tid] = dd;
        if (tid == 0) {
          s.chunk_off[T] = total;
          s.ncand = s.q_n;                                    // record space of this chunk
          s.q_n += total;
          if (s.q_n > p.adj_cap) B2K_SET_ERR(s, B2K_ERR_OVERFLOW);
        }
        __syncthreads();
        if (s.err) break;                                     // uniform
        const int ebase = s.ncand;
        if (deg > 0) {
          int4 *rp = &x.rec[dd];
          *reinterpret_cast<int2 *>(rp) = make_int2(ebase + off, deg);
          rp->z = 0;                                          // set when this expansion admits an arc
        }
        __syncthreads();
        for (int j = tid; j < total; j += T) {
          int lo = 0, hi = T;
          while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (s.chunk_off[mid] <= j) lo = mid; else hi = mid;
          }
          const int a = s.chunk_ebeg[lo] + (j - s.chunk_off[lo]);
          const int owner = s.chunk_d[lo];
          const int4 arc = __ldg(&g.ne_arcs[a]);
          const float tot = s.chunk_cost[lo] + __int_as_float(arc.y);
          int4 entry = make_int4(0, kInfBits, a, -1);
          uint32_t lowered_from = 0u;
          if (tot < cutoff) {
            bool created; int idx = 0;
            const int ds = hash_insert_b(ctx, arc.x, Hc, &created, &idx);
            if (ds >= 0) {
              const bool fill = created && idx < p.max_tpf;
              int2 o0 = make_int2(0, 0), o1 = make_int2(0, 0);
              if (fill) {                                     // loads now, the store that needs them after the atomics below
                x.rec[idx] = make_int4(0, 0, 0, 0);
                x.c0[idx] = B2K_INF_ORD;
                o0 = __ldg(&g.st_off[arc.x]); o1 = __ldg(&g.st_off[arc.x + 1]);
              }
              entry.x = __float_as_int(tot); entry.y = arc.y; entry.w = ds;
              x.rec[owner].z = 1;                             // only zero / non-zero is ever read: a plain store, no atomic round trip
              const uint32_t nv = f2ord(tot);
              const uint32_t old = atomicMin(reinterpret_cast<uint32_t *>(&hash[ds].y), nv);
              if (nv < old) {
                lowered_from = old;
                if (arc.w >= 0) {
                  if (atomicExch(&x.qstamp[ds], (uint32_t)stamp) != (uint32_t)stamp) {
                    int q = atomicAdd(&s.wl_n[cur ^ 1], 1);
                    if (q < p.max_tpf) out[q] = ds;
                    else B2K_SET_ERR(s, B2K_ERR_OVERFLOW);
                  }
                }
              }
              if (fill) x.trec[idx] = make_int4(o0.x, o1.x - o0.x, o0.y, o1.y - o0.y);
            }
          }
          x.adj[ebase + j] = entry;
          x.adjo[ebase + j] = owner;
          x.adjc[ebase + j] = lowered_from;
        }
        __syncthreads();
      }
      __syncthreads();
      if (tid == 0) {
        s.wl_n[cur] = 0;
        s.stamp = stamp;
        if (s.wl_n[cur ^ 1] > p.max_tpf) s.wl_n[cur ^ 1] = p.max_tpf;
        s.cont = (s.wl_n[cur ^ 1] > 0 && !s.err);
      }
      cur ^= 1;
      first_round = false;
      __syncthreads();
    }
  }
  __syncthreads();
  B2K_TICK(s, 5);
  const int Nall = min(s.ntok_new, p.max_tpf);
  const int E = s.err ? 0 : min(s.q_n, p.adj_cap);
  // destination slot -> creation index; the cost a lowered token had before the closure = the largest value a
  // successful atomicMin returned for it (the values a slot takes are decreasing)
  for (int e = tid; e < E; e += T) {
    const int dsl = x.adj[e].w;
    if (dsl < 0) continue;
    const int jd = hash[dsl].w;
    x.adj[e].w = jd;
    const uint32_t ov = x.adjc[e];
    if (ov) atomicMax(&x.c0[jd], ov);
  }
  __syncthreads();
  // An arc can only ever fire in the replay if final_cost(src) + w < cost(dest) after ProcessEmitting.  A destination
  // the closure never lowered keeps that cost, which is <= every admitted tot: none of its arcs can fire.
  for (int e = tid; e < E; e += T) {
    int4 en = x.adj[e];
    if (en.w < 0) continue;
    const uint32_t c0b = x.c0[en.w];
    if (!(c0b != 0u && __int_as_float(en.x) < ord2f(c0b))) x.adj[e].y = kInfBits;
    x.adj[e].x = en.w;
  }
  __syncthreads();
  if (tid == 0) cs->arcs_ne += (unsigned long long)E;        // eps arcs examined by the closure
  // ---- replay (creation order of the eps-created tokens), restricted to the ancestors of created tokens
  int *mark = wl1;                                           // the closure's worklists are idle now
  int qcarry = 0;
  extern __shared__ __align__(16) unsigned char dyn_smem_base[];
  auto cost_before_closure = [&](int d) -> float {
    const uint32_t c0b = x.c0[d];
    return c0b ? ord2f(c0b) : ord2f((uint32_t)hash[tokslot[d]].y);
  };
  bool owners_marked = false;                                // every entry left alive has a marked owner (shared-memory marking ran)
  if (Nall > N1 && !s.err) {
    if ((size_t)E * 8 + 2 * (size_t)Nall + 16 <= (size_t)p.rs_bytes) {
      owners_marked = true;
      // the fixed point runs out of shared memory (the replay arrays are not built yet): one pass reads the entries,
      // the iterations touch no scratch
      uint2 *ent_s = reinterpret_cast<uint2 *>(dyn_smem_base);      // {destination or ~0, owner | 1 << 31 if it cannot propagate}
      unsigned char *mark_s = reinterpret_cast<unsigned char *>(ent_s + E);
      unsigned char *live_s = mark_s + Nall;                        // owner of at least one entry that can fire in the replay
      for (int d = tid; d < Nall; d += T) { mark_s[d] = (d >= N1); live_s[d] = 0; if (d >= N1) x.newseq[d - N1] = -1; }
      for (int e = tid; e < E; e += T) {
        const int4 en = x.adj[e];
        uint2 v = make_uint2(0xffffffffu, 0x80000000u);
        if (en.w >= 0) {
          const int o = x.adjo[e];
          bool prop = en.y != kInfBits;
          if (prop) { const int4 r = x.rec[o]; prop = (e >= r.x && e < r.x + r.y); }   // not a superseded record
          v = make_uint2((uint32_t)en.x, (uint32_t)o | (prop ? 0u : 0x80000000u));
        }
        ent_s[e] = v;
      }
      __syncthreads();
      for (int it = 0; it < 1000000; it++) {
        if (tid == 0) s.cont = 0;
        __syncthreads();
        for (int e = tid; e < E; e += T) {
          const uint2 v = ent_s[e];
          if ((v.y & 0x80000000u) || !mark_s[v.x] || mark_s[v.y]) continue;
          mark_s[v.y] = 1;
          s.cont = 1;
        }
        __syncthreads();
        const int again = s.cont;
        __syncthreads();
        if (!again) break;
      }
      for (int d = tid; d < Nall; d += T) mark[d] = mark_s[d];
      for (int e = tid; e < E; e += T) {
        const uint2 v = ent_s[e];
        // dead for the replay: destination not an ancestor of a created token, or the entry cannot propagate (superseded
        // record / cannot fire) -- the links pass reads only arc id and destination of the final records
        if (v.x == 0xffffffffu) continue;
        if (!mark_s[v.x] || (v.y & 0x80000000u)) x.adj[e].y = kInfBits;
        else live_s[v.y] = 1;
      }
    } else {
    for (int d = tid; d < Nall; d += T) { mark[d] = (d >= N1); if (d >= N1) x.newseq[d - N1] = -1; }
    __syncthreads();
    for (int it = 0; it < 1000000; it++) {
      if (tid == 0) s.cont = 0;
      __syncthreads();
      for (int e = tid; e < E; e += T) {
        const int4 en = x.adj[e];
        if (en.w < 0 || en.y == kInfBits || !mark[en.x]) continue;
        const int o = x.adjo[e];
        if (mark[o]) continue;
        const int4 r = x.rec[o];
        if (e < r.x || e >= r.x + r.y) continue;             // superseded record
        mark[o] = 1;
        s.cont = 1;
      }
      __syncthreads();
      const int again = s.cont;
      __syncthreads();
      if (!again) break;
    }
    for (int e = tid; e < E; e += T) {
      const int4 en = x.adj[e];
      if (en.w >= 0 && !mark[en.x]) x.adj[e].y = kInfBits;
    }
    }
    __syncthreads();                                         // (mark / adj are read below; the key area reuses the shared arrays)
    // initial worklist (:852-856) = the emitting tokens in list order, restricted to the marked tokens whose final
    // record admits something.  List order = (first key of the token's bucket, own key): both are read off the table.
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(dyn_smem_base);
    const int key_cap = 1 << (31 - __clz(max(1, min(4096, p.rs_rcap + p.rs_ecap))));   // 8 bytes each, inside the walk's (not yet filled) arrays; a power of two
    if (tid == 0) s.rs_n = 0;
    __syncthreads();
    int *pend = x.big;                                        // tokens to key (idle until the walk arrays are built)
    if (owners_marked) {
      // a token none of whose entries can fire would be popped and visit no arc: left out (the order of the others is unchanged)
      const unsigned char *live_s = reinterpret_cast<const unsigned char *>(reinterpret_cast<uint2 *>(dyn_smem_base) + E) + Nall;
      for (int d = tid; d < N1; d += T)
        if (live_s[d]) pend[atomicAdd(&s.rs_n, 1)] = d;
    } else {
      for (int d = tid; d < N1; d += T) {
        if (!mark[d] || x.rec[d].z <= 0) continue;
        pend[atomicAdd(&s.rs_n, 1)] = d;
      }
    }
    __syncthreads();
    qcarry = s.rs_n;
    int P2 = 1;
    while (P2 < qcarry) P2 <<= 1;
    if (qcarry > key_cap || P2 > key_cap) keys = reinterpret_cast<unsigned long long *>(x.hw);   // (idle until the list-order pass)
    for (int q = tid; q < P2; q += T) {
      unsigned long long key = ~0ull;
      if (q < qcarry) {
        const int d = pend[q];
        const int4 hs = hash[tokslot[d]];
        const uint32_t b = bucket_b(ctx, hs.x);
        uint32_t F = (uint32_t)hs.z;
        for (int i = 0; i <= ctx.hash_mask + B2K_V2_L1_PROBES; i++) {
          const uint32_t ps = probe_slot_b(ctx, b, i);
          const int4 o = hash[ps];
          if (o.x == B2K_HASH_EMPTY) break;
          if (probe_revisits_b(ctx, b, i, ps)) continue;
          if (bucket_b(ctx, o.x) == b) F = min(F, (uint32_t)o.z);   // (eps-created tokens still carry the largest key)
        }
        key = ((unsigned long long)F << 37) | ((unsigned long long)(uint32_t)hs.z << 17) | (unsigned long long)(uint32_t)d;
      }
      keys[q] = key;
    }
    __syncthreads();
    if (qcarry <= 1024) {
      // the keys are distinct: a key's position is the number of smaller keys (one pass, no barriers inside)
      for (int i = tid; i < qcarry; i += T) {
        const unsigned long long k = keys[i];
        int r = 0;
        for (int j = 0; j < qcarry; j++) r += (keys[j] < k);
        x.queue[r] = (int)(uint32_t)(k & 0x1ffffull);
      }
    } else {
      block_bitonic_sort_u64<T>(keys, P2);
      for (int i = tid; i < qcarry; i += T) x.queue[i] = (int)(uint32_t)(keys[i] & 0x1ffffull);
    }
    __syncthreads();
  }
  __syncthreads();
  B2K_TICK(s, 6);
  // ---- compact walk out of shared memory (same construction as the first generation; falls back to the walk over the
  //      global records when the set does not fit)
  unsigned char *dyn_smem = dyn_smem_base;
  float2 *tk_s = reinterpret_cast<float2 *>(dyn_smem);        // per token {replay cost, record offset | count << 16}
  float2 *en_s = tk_s + p.rs_rcap;                            // per arc {weight, dest id}
  unsigned short *ns_s = reinterpret_cast<unsigned short *>(en_s + p.rs_ecap);
  unsigned short *q_s = ns_s + p.rs_rcap;
  int *cid = x.cid;
  int *dof = wl0;                                            // compact id -> creation index (worklists are idle now)
  bool replay_done = false;
  if (p.rs_rcap > 0 && !s.err && Nall > N1 && qcarry <= p.rs_qcap) {
    if (tid == 0) { s.rs_n = 0; s.rs_e = 0; s.rs_ok = 1; s.rs_eov = 0; }
    // compact ids of the replay's tokens: 16-bit, in shared memory behind the walk's arrays when the frame fits
    // They live in the upper half of the arc area (2 x ecap ids): the walk's arcs fit the lower half in almost every
    // frame; when they do not, the ids are spilled to scratch and the arcs are compacted again into the whole area.
    unsigned short *cid_s = reinterpret_cast<unsigned short *>(en_s + p.rs_ecap / 2);
    const bool cid_in_smem = owners_marked && p.cid_smem && p.rs_ecap >= 2 && Nall <= 2 * p.rs_ecap;
    bool use_s = cid_in_smem;
    auto cidv = [&](int d) -> int {
      if (use_s) { const unsigned v = cid_s[d]; return v >= 0xfffeu ? -1 : (int)v; }
      return cid[d];
    };
    if (cid_in_smem) {
      for (int d = tid; d < Nall; d += T) cid_s[d] = 0xffffu;
      __syncthreads();
      for (int k = tid; k < qcarry; k += T) cid_s[x.queue[k]] = 0xfffeu;
      for (int e = tid; e < E; e += T) {
        const int4 en = x.adj[e];
        if (en.w < 0 || en.y == kInfBits) continue;
        cid_s[en.x] = 0xfffeu;
        cid_s[x.adjo[e]] = 0xfffeu;
      }
      __syncthreads();
      for (int d = tid; d < Nall; d += T) {
        if (cid_s[d] != 0xfffeu) continue;
        const int id = atomicAdd(&s.rs_n, 1);
        if (id < p.rs_rcap) {
          tk_s[id] = make_float2(cost_before_closure(d), __int_as_float(0));
          ns_s[id] = 0xffff;
          dof[id] = d;
          cid_s[d] = (unsigned short)id;
          if (d >= N1) x.newseq[d - N1] = id;                 // (kept here: the id area is reused by the parallel walk)
        } else {
          s.rs_ok = 0;
        }
      }
    } else {
    for (int d = tid; d < Nall; d += T) cid[d] = -1;
    __syncthreads();
    auto claim = [&](int d) {
      if (atomicCAS(&cid[d], -1, -2) != -1) return;
      int id = atomicAdd(&s.rs_n, 1);
      if (id < p.rs_rcap) {
        tk_s[id] = make_float2(cost_before_closure(d), __int_as_float(0));
        ns_s[id] = 0xffff;
        dof[id] = d;
        cid[d] = id;
      } else {
        s.rs_ok = 0;
      }
    };
    for (int k = tid; k < qcarry; k += T) claim(x.queue[k]);
    for (int e = tid; e < E; e += T) {
      int4 en = x.adj[e];
      if (en.w < 0 || en.y == kInfBits) continue;
      int o = x.adjo[e];
      if (!mark[o]) continue;
      if (cid[en.x] == -1) claim(en.x);
      if (cid[o] == -1) claim(o);
    }
    }
    __syncthreads();
    for (int attempt = 0; attempt < 2; attempt++) {
    const int ecap_eff = use_s ? p.rs_ecap / 2 : p.rs_ecap;
    if (s.rs_ok) {
      const int R = s.rs_n;
      const int lane_id = tid & 31;
      if (tid == 0) s.ncand = 0;
      __syncthreads();
      int *big_list = x.big;
      for (int id = tid; id < R; id += T) {
        const int d = dof[id];
        const int4 r = x.rec[d];
        if (!mark[d] || r.z == 0) continue;                  // destination only
        if (r.y > 8) { big_list[atomicAdd(&s.ncand, 1)] = id; continue; }
        int4 en[8];
        int total = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          en[i] = (i < r.y) ? x.adj[r.x + i] : make_int4(0, kInfBits, 0, -1);
          total += (en[i].w >= 0 && en[i].y != kInfBits);
        }
        if (!total) continue;
        const int eo = atomicAdd(&s.rs_e, total);
        if (eo + total > ecap_eff) { s.rs_ok = 0; s.rs_eov = 1; continue; }
        int pos = eo;
#pragma unroll
        for (int i = 0; i < 8; i++)
          if (en[i].w >= 0 && en[i].y != kInfBits)
            en_s[pos++] = make_float2(__int_as_float(en[i].y), __int_as_float(cidv(en[i].x)));
        tk_s[id].y = __int_as_float(eo | (total << 16));
      }
      __syncthreads();
      const int nbig = s.ncand;
      for (int bi = tid >> 5; bi < nbig; bi += T / 32) {     // one warp per hub record
        const int id = big_list[bi];
        const int d = dof[id];
        const int4 r = x.rec[d];
        int total = 0;
        for (int b0 = 0; b0 < r.y; b0 += 32) {
          int i = b0 + lane_id;
          int4 en = (i < r.y) ? x.adj[r.x + i] : make_int4(0, kInfBits, 0, -1);
          total += __popc(__ballot_sync(0xffffffffu, en.w >= 0 && en.y != kInfBits));
        }
        if (!total) continue;
        int eo = 0;
        if (lane_id == 0) eo = atomicAdd(&s.rs_e, total);
        eo = __shfl_sync(0xffffffffu, eo, 0);
        if (eo + total > ecap_eff || total > 0xffff) { if (lane_id == 0) { s.rs_ok = 0; if (total <= 0xffff) s.rs_eov = 1; } continue; }
        int pos = eo;
        for (int b0 = 0; b0 < r.y; b0 += 32) {
          int i = b0 + lane_id;
          int4 en = (i < r.y) ? x.adj[r.x + i] : make_int4(0, kInfBits, 0, -1);
          bool keep = en.w >= 0 && en.y != kInfBits;
          uint32_t m = __ballot_sync(0xffffffffu, keep);
          if (keep) {
            int q = pos + __popc(m & ((1u << lane_id) - 1u));
            en_s[q] = make_float2(__int_as_float(en.y), __int_as_float(cidv(en.x)));
          }
          pos += __popc(m);
        }
        if (lane_id == 0) tk_s[id].y = __int_as_float(eo | (total << 16));
      }
      for (int k = tid; k < qcarry; k += T) q_s[k] = (unsigned short)cidv(x.queue[k]);
    }
    __syncthreads();
    if (!(use_s && s.rs_eov && attempt == 0)) break;         // uniform
    // the arcs did not fit beside the ids: ids to scratch, arcs again into the whole area
    for (int d = tid; d < Nall; d += T) cid[d] = cidv(d);
    for (int id = tid; id < min(s.rs_n, p.rs_rcap); id += T) tk_s[id].y = __int_as_float(0);
    __syncthreads();
    if (tid == 0) { s.rs_e = 0; s.rs_ok = 1; s.rs_eov = 0; }
    use_s = false;
    __syncthreads();
    }
    // ---- An initial entry none of whose arcs fires under the INITIAL costs never does anything when it is popped: either its
    //      token's cost is still the initial one then (destinations have only become cheaper), or it was lowered before,
    //      in which case the token was pushed and processed on the spot and every arc was relaxed with the lowered cost.
    //      Such entries are dropped (in parallel; the order of the others is kept).
    int qinit = qcarry;                                      // entries of the shared-memory walks (the scratch walk keeps the full list)
    if (s.rs_ok && qcarry > 0) {
      int kept_total = 0;
      for (int base = 0; base < qcarry; base += T) {
        const int k = base + tid;
        int keep = 0, d = 0;
        if (k < qcarry) {
          d = q_s[k];
          const float2 td = tk_s[d];
          if (td.x < cutoff) {
            const int oc = __float_as_int(td.y);
            const int e0 = oc & 0xffff, e1 = e0 + (int)((unsigned)oc >> 16);
            for (int e = e0; e < e1 && !keep; e++) {
              const float2 ent = en_s[e];
              const float tot = td.x + ent.x;
              keep = (tot < cutoff) && (tot < tk_s[__float_as_int(ent.y)].x);
            }
          }
        }
        int total;
        const int off = block_excl_scan<T>(keep, s.redi, &total);   // (barriers inside: every thread has read its q_s[k] of this chunk)
        if (keep) q_s[kept_total + off] = (unsigned short)d;         // kept_total + off <= k: never ahead of an unread entry of a later chunk
        kept_total += total;
        __syncthreads();
      }
      qinit = kept_total;
    }
    // ---- The replay is a stack: the initial entries are popped from the back and everything an entry pushes is processed
    //      before the entry below it, so the run is a sequence of cascades, one per initial entry.  Cascades only interact
    //      through tokens they share: over the connected components of the compact graph they are independent, and the
    //      creation order of the new tokens is (order of the cascade's initial entry, order inside the cascade).  So: label
    //      the components (min-label propagation with pointer jumping, all in shared memory), one thread per component
    //      replays its entries in stack order with a private stack, every created token is recorded with that key, and
    //      the keys are ranked.  (One thread for everything was 22 % of the frame: ~470 pops at ~600 cycles each.)
    bool par_done = false;
    if (use_s && s.rs_ok && p.par_walk && min(s.rs_n, p.rs_rcap) <= p.rs_ecap) {     // uniform
      const int R = min(s.rs_n, p.rs_rcap);
      int *lab = reinterpret_cast<int *>(cid_s);              // R ints in the id area (ids of created tokens are in x.newseq)
      for (int id = tid; id < R; id += T) lab[id] = id;
      if (tid == 0) s.ncand = 0;
      __syncthreads();
      for (int it = 0; it < 100000; it++) {
        if (tid == 0) s.cont = 0;
        __syncthreads();
        for (int id = tid; id < R; id += T) {
          const int oc = __float_as_int(tk_s[id].y);
          const int e0 = oc & 0xffff, e1 = e0 + (int)((unsigned)oc >> 16);
          const int mine = lab[id];
          int m = min(mine, lab[mine]);                       // pointer jumping
          for (int e = e0; e < e1; e++) m = min(m, lab[__float_as_int(en_s[e].y)]);
          bool ch = false;
          if (m < mine) { atomicMin(&lab[id], m); ch = true; }
          for (int e = e0; e < e1; e++) {
            const int j = __float_as_int(en_s[e].y);
            if (m < lab[j]) { atomicMin(&lab[j], m); ch = true; }
          }
          if (ch) s.cont = 1;
        }
        __syncthreads();
        const int again = s.cont;
        __syncthreads();
        if (!again) break;
      }
      B2K_TICK(s, 3);
      unsigned long long *crec = reinterpret_cast<unsigned long long *>((reinterpret_cast<uintptr_t>(x.queue + qcarry) + 15) & ~(uintptr_t)15);   // behind the initial worklist (kept for the fallback)
      for (int root = tid; root < R; root += T) {
        if (lab[root] != root) continue;
        unsigned short stk[32];
        bool overflow = false;
        for (int k = qinit - 1; k >= 0 && !overflow; k--) {
          const int d0 = q_s[k];
          if (lab[d0] != root) continue;
          int sp = 0, seq = 0;
          stk[sp++] = (unsigned short)d0;
          while (sp > 0) {
            const int d = stk[--sp];
            const float2 td = tk_s[d];
            const float c = td.x;
            if (c >= cutoff) continue;
            const int oc = __float_as_int(td.y);
            const int e0 = oc & 0xffff, e1 = e0 + (int)((unsigned)oc >> 16);
            for (int e = e0; e < e1; e++) {
              const float2 ent = en_s[e];
              const float tot = c + ent.x;
              if (tot < cutoff) {
                const int j = __float_as_int(ent.y);
                const float2 tj = tk_s[j];
                if (tot < tj.x) {
                  tk_s[j].x = tot;
                  if (tj.x == kInfF) {
                    const int slot = atomicAdd(&s.ncand, 1);
                    crec[slot] = ((unsigned long long)(qinit - 1 - k) << 32) | ((unsigned long long)seq++ << 12) | (unsigned long long)j;
                  }
                  if ((unsigned)__float_as_int(tj.y) >> 16) {
                    if (sp < 32) stk[sp++] = (unsigned short)j;
                    else { overflow = true; sp = 0; break; }
                  }
                }
              }
            }
          }
        }
        if (overflow) s.rs_ok = 0;                             // (the walk over the scratch records redoes the frame)
      }
      __syncthreads();
      if (s.rs_ok) {
        const int n = s.ncand;
        if (n != Nall - N1) { if (tid == 0) B2K_SET_ERR(s, B2K_ERR_STATE); }
        else if (n > 4096) { if (tid == 0) s.rs_ok = 0; }
        else {
          for (int i = tid; i < n; i += T) {                  // the records are distinct: rank = number of smaller ones
            const unsigned long long mine = crec[i];
            int r = 0;
            for (int j2 = 0; j2 < n; j2++) r += (crec[j2] < mine);
            ns_s[(int)(mine & 0xfffull)] = (unsigned short)r;
          }
          if (tid == 0) s.prof[14] += 1;
        }
      }
      __syncthreads();
      par_done = s.rs_ok != 0;
      __syncthreads();
    } else {
      B2K_TICK(s, 3);                                        // (profile: slot 3 = building the compact arrays, slot 7 = the walk)
    }
    // (profile: which frames take which route -- three 20-bit counters: marking in shared memory, ids in shared memory, walk by components)
    if (PROF && tid == 0) s.prof[2] += (unsigned long long)owners_marked | ((unsigned long long)use_s << 20) | ((unsigned long long)par_done << 40);
    if (!par_done && s.rs_ok && tid == 0) {
      int qn = qinit, next = 0;
      bool ok = true;
      int npop = 0, nvis = 0;
      while (qn > 0) {
        const int d = q_s[--qn];
        const float2 td = tk_s[d];
        const float c = td.x;
        npop++;
        if (c >= cutoff) continue;
        const int oc = __float_as_int(td.y);
        const int e0 = oc & 0xffff, e1 = e0 + (int)((unsigned)oc >> 16);
        nvis += e1 - e0;
        for (int e = e0; e < e1; e++) {
          const float2 ent = en_s[e];
          const float tot = c + ent.x;
          if (tot < cutoff) {
            const int j = __float_as_int(ent.y);
            const float2 tj = tk_s[j];
            if (tot < tj.x) {
              tk_s[j].x = tot;
              if (tj.x == kInfF) ns_s[j] = (unsigned short)next++;
              if ((unsigned)__float_as_int(tj.y) >> 16) {
                if (qn < p.rs_qcap) q_s[qn++] = (unsigned short)j;
                else { ok = false; qn = 0; break; }
              }
            }
          }
        }
      }
      if (ok) {
        if (next != Nall - N1) B2K_SET_ERR(s, B2K_ERR_STATE);
        s.prof[12] += (unsigned long long)npop; s.prof[13] += (unsigned long long)nvis; s.prof[14] += 1;
      } else {
        s.rs_ok = 0;                                         // worklist outgrew shared memory: redo below
      }
    }
    __syncthreads();
    if (s.rs_ok) {
      replay_done = true;
      if (!s.err)
        for (int d = N1 + tid; d < Nall; d += T) {
          int id = use_s ? x.newseq[d - N1] : cidv(d);        // (shared-memory ids: stored at claim time, the area may have been reused)
          if (id >= 0) x.newseq[d - N1] = (int)ns_s[id];
          else B2K_SET_ERR(s, B2K_ERR_STATE);                // an eps-created token is always some record's destination
        }
    }
    __syncthreads();
  }
  // literal LIFO replay by one thread over the dense records in global memory (rec.w = replay cost)
  if (!s.err && !replay_done && Nall > N1) {
    for (int d = tid; d < Nall; d += T) x.rec[d].w = __float_as_int(cost_before_closure(d));
    __syncthreads();
    if (tid == 0) {
      int qn = qcarry, next = 0;
      int npop = 0, nvis = 0;
      while (qn > 0) {
        const int d = x.queue[--qn];
        const int4 r = x.rec[d];
        const float c = __int_as_float(r.w);
        npop++;
        if (c >= cutoff || r.z == 0) continue;
        nvis += r.y;
        for (int e = r.x; e < r.x + r.y; e++) {
          const int4 en = x.adj[e];
          if (en.w < 0) continue;
          const float tot = c + __int_as_float(en.y);
          if (tot < cutoff) {
            const int j = en.x;
            const int4 rj = x.rec[j];
            const float old = __int_as_float(rj.w);
            if (tot < old) {                                   // FindOrAddToken: new or improved -> changed
              x.rec[j].w = __float_as_int(tot);
              if (old == kInfF) x.newseq[j - N1] = next++;
              if (rj.z > 0) {
                if (qn < p.queue_cap) x.queue[qn++] = j;
                else { B2K_SET_ERR(s, B2K_ERR_OVERFLOW); break; }
              }
            }
          }
        }
      }
      if (!s.err && next != Nall - N1) B2K_SET_ERR(s, B2K_ERR_STATE);
      s.prof[12] += (unsigned long long)npop; s.prof[13] += (unsigned long long)nvis;
    }
  }
  __syncthreads();
  B2K_TICK(s, 7);
  if (s.err) return;                                         // uniform (the caller resets the table)
  // ---- insertion keys of the eps-created tokens: after every arc position, in replay order
  const int PP = P + (Nall - N1);
  for (int d = N1 + tid; d < Nall; d += T) hash[tokslot[d]].z = P + x.newseq[d - N1];
  // the key space fits the (now idle) replay arrays in most frames: bucket heads, their scan and the commit's look-ups
  // stay in shared memory instead of three round trips to scratch that misses L2
  uint32_t *hwp = ((size_t)PP * 4 <= (size_t)p.rs_bytes) ? reinterpret_cast<uint32_t *>(dyn_smem_base) : x.hw;
  for (int w = tid; w < PP; w += T) hwp[w] = 0u;
  __syncthreads();
  B2K_TICK(s, 8);
  // ---- HashList order (hash-list-inl.h:126-175): buckets by first insertion, insertion order inside a bucket.  Every
  //      token walks its bucket's probe run: first key of the bucket, its own rank inside it, the population.
  //      (A scan of the whole level-1 window with the runs analysed per thread was measured slower, 578 vs 533 ms:
  //      profiles/r02_decoder_history.md.)
  for (int d = tid; d < Nall; d += T) {
    const int sl = tokslot[d];
    const int4 hs = hash[sl];
    const uint32_t b = bucket_b(ctx, hs.x);
    uint32_t F = (uint32_t)hs.z;
    int within = 0, pop = 0;
    for (int i = 0; i <= ctx.hash_mask + B2K_V2_L1_PROBES; i++) {
      const uint32_t ps = probe_slot_b(ctx, b, i);
      const int4 o = hash[ps];
      if (o.x == B2K_HASH_EMPTY) break;
      if (probe_revisits_b(ctx, b, i, ps)) continue;
      if (bucket_b(ctx, o.x) == b) {
        pop++;
        F = min(F, (uint32_t)o.z);
        within += ((uint32_t)o.z < (uint32_t)hs.z);
      }
    }
    x.tok4[d] = make_int4(hs.x, __float_as_int(ord2f((uint32_t)hs.y)), (int)F, within);
    if (F == (uint32_t)hs.z) hwp[F] = (uint32_t)pop;
  }
  __syncthreads();
  block_excl_scan_array<T>(hwp, PP, s.redi);
  B2K_TICK(s, 9);
  const bool fits = (ctx.tbase + Nall <= p.max_tokens);
  if (!fits) { if (tid == 0) B2K_SET_ERR(s, B2K_ERR_OVERFLOW); __syncthreads(); return; }
  // ---- commit: tokens to the arena in list order, the emitting ranges for the next frame's expansion
  for (int d = tid; d < Nall; d += T) {
    const int4 t = x.tok4[d];
    const int r = (int)hwp[t.z] + t.w;
    tok_state[ctx.tbase + r] = t.x;
    tok_cost[ctx.tbase + r] = __int_as_float(t.y);
    const int4 tr = x.trec[d];
    x.pf_ebeg[r] = tr.x;
    x.pf_edeg[r] = tr.y;
    x.rank[d] = r;
  }
  __syncthreads();
  // ---- eps links = the admitted entries of the final records, written with arena indices
  {
    const int lane_id = tid & 31;
    for (int base = 0; base < E; base += T) {
      const int e = base + tid;
      bool live = false;
      int4 en = make_int4(0, 0, 0, -1);
      int o = 0;
      if (e < E) {
        en = x.adj[e];
        if (en.w >= 0) {
          o = x.adjo[e];
          const int4 r = x.rec[o];
          live = (e >= r.x && e < r.x + r.y);
        }
      }
      const uint32_t m = __ballot_sync(0xffffffffu, live);
      if (m) {
        int lb = 0;
        if (lane_id == 0) lb = atomicAdd(&s.nlink_new, __popc(m));
        lb = __shfl_sync(0xffffffffu, lb, 0);
        if (live) {
          int li = lb + __popc(m & ((1u << lane_id) - 1u));
          if (lbase + li < p.max_links)
            links[lbase + li] = make_int4(ctx.tbase + x.rank[o], ctx.tbase + x.rank[en.w], (int)((uint32_t)en.z | B2K_EPS_FLAG), 0);
          else
            B2K_SET_ERR(s, B2K_ERR_OVERFLOW);
        }
      }
    }
  }
  // emitting links: destination slot -> arena index
  for (int li = tid; li < n_emit_links; li += T) {
    int4 *lp = &links[lbase + li];
    lp->y = ctx.tbase + x.rank[hash[lp->y].w];
  }
  __syncthreads();
  B2K_TICK(s, 10);
  const int nlink = s.nlink_new;
  for (int d = tid; d < Nall; d += T)
    hash[tokslot[d]] = make_int4(B2K_HASH_EMPTY, (int)B2K_INF_ORD, 0x7fffffff, 0);
  if (tid == 0) {
    size_t fo = (size_t)ch * (p.max_frames + 2);
    p.frame_tok_begin[fo + list_index] = ctx.tbase;
    p.frame_tok_begin[fo + list_index + 1] = ctx.tbase + Nall;
    p.frame_link_begin[fo + list_index] = lbase;
    p.frame_link_eps[fo + list_index] = lbase + n_emit_links;
    p.frame_link_begin[fo + list_index + 1] = lbase + nlink;
    if (list_index > 0) {
      size_t co = (size_t)ch * (p.max_frames + 1) + (list_index - 1);
      p.frame_cost_offset[co] = cost_offset;
      p.frame_cutoff[co] = cutoff;
    }
    B2K_TICK(s, 11);
  }
  __syncthreads();
}

template <int T, bool PROF, int IT>
__device__ void dec_advance_v2_lane(const DecParams &p, DecShared<T> &s, const int lane, const int slot) {
  const int tid = threadIdx.x;
  const int ch = p.lane_channel[lane];
  ChanState *cs = &p.chan[ch];
  const FstDev &g = p.fst;
  const float kInf = __int_as_float(0x7f800000);

  if (cs->status != B2K_OK) return;
  if (!p.do_init && cs->frames_decoded < 0) {
    if (tid == 0) { cs->status = B2K_ERR_STATE; cs->err_line = __LINE__; }
    return;
  }
  int32_t *tok_state = p.tok_state + (size_t)ch * p.max_tokens;
  float *tok_cost = p.tok_cost + (size_t)ch * p.max_tokens;
  int4 *links = p.links + (size_t)ch * p.max_links;

  LaneCtx ctx;
  ctx.hash = p.hash + (size_t)slot * p.hash_size;
  ctx.tokslot = p.tokslot + (size_t)slot * p.max_tpf;
  ctx.tok_state = tok_state;
  ctx.hash_mask = p.hash_size - 1;
  ctx.hash_log = p.hash_log;
  ctx.max_tpf = p.max_tpf;
  ctx.max_tokens = p.max_tokens;
  ctx.ntok_new = &s.ntok_new;
  ctx.err = &s.err;
  ctx.err_line = &s.err_line;
  set_l1_b(ctx, 1000u, p.v2_l1_shift);
  ctx.used_l2 = &s.used_l2;
  X2 x;
  x.trec = p.v2_trec + (size_t)slot * p.max_tpf;
  x.rec = p.x_rec + (size_t)slot * p.max_tpf;
  x.c0 = p.v2_c0 + (size_t)slot * p.max_tpf;
  x.newseq = p.x_newseq + (size_t)slot * p.max_tpf;
  x.adj = p.x_adj + (size_t)slot * p.adj_cap;
  x.adjo = p.x_adjo + (size_t)slot * p.adj_cap;
  x.adjc = p.v2_adjc + (size_t)slot * p.adj_cap;
  x.wl0 = p.wl + (size_t)slot * 2 * p.max_tpf;
  x.wl1 = x.wl0 + p.max_tpf;
  x.queue = p.cand + (size_t)slot * 5 * p.cand_cap;
  x.qstamp = p.v2_qstamp + (size_t)slot * p.hash_size;
  x.tok4 = p.v2_tok4 + (size_t)slot * p.max_tpf;
  x.hw = p.v2_hw + (size_t)slot * p.v2_hw_len;
  x.rank = p.x_order + (size_t)slot * p.max_tpf;
  x.pf_ebeg = p.v2_pf + (size_t)slot * 2 * p.max_tpf;
  x.pf_edeg = x.pf_ebeg + p.max_tpf;
  x.cid = p.x_sbase + (size_t)slot * p.max_tpf;
  x.big = p.x_run + (size_t)slot * p.max_tpf;

  if (tid == 0) { s.err = 0; s.err_line = 0; s.stamp = p.lane_stamp[slot]; }
  __syncthreads();

  if (p.do_init) {
    if (tid == 0) { s.ntok_new = 0; s.nlink_new = 0; s.used_l2 = 0; }
    __syncthreads();
    ctx.tbase = 0;
    if (tid == 0) {
      bool created; int idx = 0;
      int sl = hash_insert_b(ctx, g.start, 1000u, &created, &idx);
      if (sl >= 0) {
        ctx.hash[sl].y = (int)f2ord(0.0f); ctx.hash[sl].z = 0;
        const int2 o0 = __ldg(&g.st_off[g.start]), o1 = __ldg(&g.st_off[g.start + 1]);
        x.trec[0] = make_int4(o0.x, o1.x - o0.x, o0.y, o1.y - o0.y);
      }
    }
    finish_frame_v2<T, PROF>(p, s, ctx, x, ch, 0, p.beam, 0.0f, 0, 1000u, 1, cs);
    if (tid == 0) {
      cs->frames_decoded = 0;
      cs->ntok = min(s.ntok_new, p.max_tpf);
      cs->nlink = s.nlink_new;
      cs->finalized = 0;
      cs->hc = 1000;                                   // toks_.SetSize(1000) (:39)
      if (s.err) { cs->status = s.err; cs->err_line = s.err_line; }
      p.lane_stamp[slot] = s.stamp;
    }
    if (s.err) reset_lane_hash<T>(ctx.hash, p.hash_size);
    return;
  }

  const int nframes = p.lane_nframes[lane];
  const float *ll_base = p.lane_loglikes[lane];
  int frames_decoded = cs->frames_decoded;
  int32_t tbase = cs->ntok, lbase = cs->nlink;
  int Hc = cs->hc;
  unsigned long long arcs_e_total = 0;
  const size_t fo = (size_t)ch * (p.max_frames + 2);

  if (tid == 0) { for (int i = 0; i < 16; i++) s.prof[i] = 0; s.tlast = PROF ? clock64() : 0; }
  const long long t_kernel0 = PROF ? clock64() : 0;
  extern __shared__ __align__(16) unsigned char dyn_smem_base[];
  float *ll_s = reinterpret_cast<float *>(dyn_smem_base);    // shares the replay arrays' space: they are idle during the expansion
  bool have_pf = false;                                      // x.pf_* describe the previous list (true after one frame of this launch)
  for (int fi = 0; fi < nframes; fi++) {
    if (frames_decoded >= p.max_frames) { if (tid == 0) B2K_SET_ERR(s, B2K_ERR_OVERFLOW); __syncthreads(); break; }
    const float *llg = ll_base + (size_t)fi * p.row_stride;
    if (p.ll_smem) for (int i = tid; i < p.num_pdfs; i += T) ll_s[i] = __ldcs(&llg[i]);
    const int pb = p.frame_tok_begin[fo + frames_decoded];
    const int pe = p.frame_tok_begin[fo + frames_decoded + 1];
    const int K = pe - pb;

    // ---- GetCutoff (:653-720); ties -> first in list order (strict < at :662/:675)
    unsigned long long local = ~0ull;
    for (int i = pb + tid; i < pe; i += T) {
      unsigned long long key = ((unsigned long long)f2ord(tok_cost[i]) << 32) | (uint32_t)(i - pb);
      local = key < local ? key : local;
    }
    unsigned long long bestkey = block_min_u64<T>(local, s.red64);
    float best_cost = kInf;
    int best_i = -1;
    if (K > 0) {
      best_cost = ord2f((uint32_t)(bestkey >> 32));
      best_i = (int)(uint32_t)(bestkey & 0xffffffffu);
    }
    const float beam_cutoff = best_cost + p.beam;
    float cur_cutoff = beam_cutoff, adaptive_beam = p.beam;
    if (K > 0 && !(p.max_active == 0x7fffffff && p.min_active == 0)) {
      int c_lt = 0, c_le = 0;
      for (int i = pb + tid; i < pe; i += T) {
        float c = tok_cost[i];
        c_lt += (c < beam_cutoff);
        c_le += (c <= beam_cutoff);
      }
      c_lt = block_sum_i32<T>(c_lt, s.redi);
      c_le = block_sum_i32<T>(c_le, s.redi);
      bool done = false;
      if (K > p.max_active && c_lt > p.max_active) {
        float mac = block_select_kth<T>(tok_cost + pb, K, p.max_active, s.hist, s.pref);
        cur_cutoff = mac;
        adaptive_beam = mac - best_cost + p.beam_delta;
        done = true;
      }
      if (!done) {
        float min_active_cutoff = kInf;
        if (K > p.min_active) {
          if (p.min_active == 0) min_active_cutoff = best_cost;
          else if (c_le <= p.min_active)
            min_active_cutoff = block_select_kth<T>(tok_cost + pb, K, p.min_active, s.hist, s.pref);
          else min_active_cutoff = beam_cutoff;
        }
        if (min_active_cutoff > beam_cutoff) {
          adaptive_beam = min_active_cutoff - best_cost + p.beam_delta;
          cur_cutoff = min_active_cutoff;
        }
      }
    }
    // PossiblyResizeHash(tok_cnt) (:227-233)
    {
      long long new_sz = (long long)((float)K * p.hash_ratio);
      if (new_sz > (long long)Hc) Hc = (int)min(new_sz, (long long)0x7fffffff);
    }
    set_l1_b(ctx, (uint32_t)Hc, p.v2_l1_shift);
    const float cost_offset = (K > 0) ? -best_cost : 0.0f;

    // ---- seed (:753-768)
    uint32_t seed_local = B2K_INF_ORD;
    if (K > 0) {
      int sb, se;
      if (have_pf) { sb = x.pf_ebeg[best_i]; se = sb + x.pf_edeg[best_i]; }
      else { const int bs = tok_state[pb + best_i]; sb = __ldg(&g.st_off[bs]).x; se = __ldg(&g.st_off[bs + 1]).x; }
      for (int a = sb + tid; a < se; a += T) {
        int4 arc = __ldg(&g.e_arcs[a]);
        float new_weight = __int_as_float(arc.y) + cost_offset - (p.ll_smem ? ll_s[arc.z] : __ldg(&llg[arc.z])) + best_cost;
        seed_local = min(seed_local, f2ord(new_weight + adaptive_beam));
      }
    }
    const float seed_cutoff = ord2f(block_min_u32<T>(seed_local, s.red32));

    // ---- main loop (:779-812): admission against the exclusive prefix-min
    if (tid == 0) { s.ntok_new = 0; s.nlink_new = 0; s.used_l2 = 0; }
    __syncthreads();
    B2K_TICK(s, 0);
    ctx.tbase = tbase;
    float carry = kInf;
    int pos_base = 0;
    int round_no = 0;
    constexpr int TPT = DecShared<T>::TPT, CT = DecShared<T>::CT;
    for (int cb = pb; cb < pe; cb += CT) {
#pragma unroll
      for (int u = 0; u < TPT; u++) {                       // striped (coalesced) loads
        const int q = u * T + tid, i = cb + q;
        int deg = 0, ebeg = 0;
        float c = 0.f;
        if (i < pe) {
          c = tok_cost[i];
          if (c <= cur_cutoff) {
            if (have_pf) { ebeg = x.pf_ebeg[i - pb]; deg = x.pf_edeg[i - pb]; }
            else {
              int st = tok_state[i];
              int2 o0 = __ldg(&g.st_off[st]), o1 = __ldg(&g.st_off[st + 1]);
              ebeg = o0.x;
              deg = o1.x - o0.x;
            }
          }
        }
        s.chunk_off[q] = deg; s.chunk_ebeg[q] = ebeg; s.chunk_cost[q] = c;
      }
      __syncthreads();
      int total;
      {                                                      // blocked exclusive scan of the degrees
        int dloc[TPT], sum = 0;
#pragma unroll
        for (int u = 0; u < TPT; u++) { dloc[u] = s.chunk_off[tid * TPT + u]; sum += dloc[u]; }
        int off = block_excl_scan<T>(sum, s.redi, &total);
#pragma unroll
        for (int u = 0; u < TPT; u++) { s.chunk_off[tid * TPT + u] = off; off += dloc[u]; }
      }
      if (tid == 0) { s.chunk_off[CT] = total; arcs_e_total += (unsigned long long)total; }
      __syncthreads();
      for (int r0 = 0; r0 < total; r0 += T * IT, round_no++) {
        const int j0 = r0 + tid * IT;
        float tots[IT], acs[IT];
        int arcid[IT], nexts[IT], srcs[IT];
        int lo = 0;
        if (j0 < total) {
          int hi = CT;
          while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (s.chunk_off[mid] <= j0) lo = mid; else hi = mid;
          }
        }
#pragma unroll
        for (int k = 0; k < IT; k++) {
          int j = j0 + k;
          tots[k] = kInf; acs[k] = 0.f; arcid[k] = 0; nexts[k] = 0; srcs[k] = 0;
          if (j < total) {
            while (s.chunk_off[lo + 1] <= j) lo++;
            int a = s.chunk_ebeg[lo] + (j - s.chunk_off[lo]);
            int4 arc = __ldg(&g.e_arcs[a]);
            float ac = cost_offset - (p.ll_smem ? ll_s[arc.z] : __ldg(&llg[arc.z]));
            tots[k] = s.chunk_cost[lo] + ac + __int_as_float(arc.y);
            acs[k] = ac; arcid[k] = a; nexts[k] = arc.x; srcs[k] = cb + lo;
            // the destination's first probe slot is on its way to L2 while the admission scan runs (B2K_DEC_PREFETCH=0 for A/B)
            if (p.hash_prefetch) asm volatile("prefetch.global.L2 [%0];" :: "l"(&ctx.hash[probe_slot_b(ctx, bucket_b(ctx, arc.x), 0)]));
          }
        }
        float ex[IT];
        float pm = kInf;
#pragma unroll
        for (int k = 0; k < IT; k++) { ex[k] = pm; pm = fminf(pm, tots[k]); }
        // block-wide exclusive scan (min) of pm in thread order
        const int lane_id = tid & 31, warp = tid >> 5;
        float incl = pm;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          float n = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane_id >= o) incl = fminf(incl, n);
        }
        float wexcl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane_id == 0) wexcl = kInf;
        float *buf = s.scanf_[round_no & 1];
        if (lane_id == 31) buf[warp] = incl;
        __syncthreads();
        float wpre = kInf, tot_all = kInf;
#pragma unroll
        for (int w = 0; w < T / 32; w++) {
          float v = buf[w];
          if (w < warp) wpre = fminf(wpre, v);
          tot_all = fminf(tot_all, v);
        }
        const float base = fminf(carry, fminf(wpre, wexcl));
        int cr[IT];                                           // creation index of a token this thread created, else -1
#pragma unroll
        for (int k = 0; k < IT; k++) {
          int j = j0 + k;
          float excl = fminf(base, ex[k]);
          float rc = fminf(seed_cutoff, excl + adaptive_beam);
          bool adm = (j < total) && (tots[k] < rc);
          int sl = -1;
          cr[k] = -1;
          if (adm) {
            bool created; int idx;
            sl = hash_insert_b(ctx, nexts[k], (uint32_t)Hc, &created, &idx);
            if (sl >= 0) {
              atomicMin(reinterpret_cast<uint32_t *>(&ctx.hash[sl].y), f2ord(tots[k]));
              atomicMin(&ctx.hash[sl].z, pos_base + j);
              if (created && idx < p.max_tpf) cr[k] = idx;
            } else adm = false;
          }
          uint32_t m = __ballot_sync(0xffffffffu, adm);
          if (m) {
            int lb = 0;
            if (lane_id == 0) lb = atomicAdd(&s.nlink_new, __popc(m));
            lb = __shfl_sync(0xffffffffu, lb, 0);
            if (adm) {
              int li = lb + __popc(m & ((1u << lane_id) - 1u));
              if (lbase + li < p.max_links)
                links[lbase + li] = make_int4(srcs[k], sl, arcid[k], __float_as_int(acs[k]));
              else
                B2K_SET_ERR(s, B2K_ERR_OVERFLOW);
            }
          }
        }
        // the creator of a token reads the state table for it, once per token and off the admission path
#pragma unroll
        for (int k = 0; k < IT; k++)
          if (cr[k] >= 0) {
            const int2 o0 = __ldg(&g.st_off[nexts[k]]), o1 = __ldg(&g.st_off[nexts[k] + 1]);
            x.trec[cr[k]] = make_int4(o0.x, o1.x - o0.x, o0.y, o1.y - o0.y);
          }
        carry = fminf(carry, tot_all);
      }
      pos_base += total;
      __syncthreads();
    }
    const float next_cutoff = fminf(seed_cutoff, carry + adaptive_beam);
    if (pos_base > p.pos_cap) { if (tid == 0) B2K_SET_ERR(s, B2K_ERR_OVERFLOW); }
    __syncthreads();
    B2K_TICK(s, 1);
    finish_frame_v2<T, PROF>(p, s, ctx, x, ch, frames_decoded + 1, next_cutoff, cost_offset, lbase, (uint32_t)Hc,
                             pos_base, cs);
    if (s.err) break;
    tbase += min(s.ntok_new, p.max_tpf);
    lbase += s.nlink_new;
    frames_decoded++;
    have_pf = true;
    __syncthreads();
  }
  if (tid == 0) {
    cs->frames_decoded = frames_decoded;
    cs->ntok = tbase;
    cs->nlink = lbase;
    cs->arcs_e += arcs_e_total;
    cs->hc = Hc;
    if (s.err) { cs->status = s.err; cs->err_line = s.err_line; }
    for (int i = 0; i < 15; i++) cs->prof[i] += s.prof[i];
    if (PROF) cs->prof[15] += (unsigned long long)(clock64() - t_kernel0);
    p.lane_stamp[slot] = s.stamp;
  }
  if (s.err) reset_lane_hash<T>(ctx.hash, p.hash_size);
}

template <int T, bool PROF, int IT>
__global__ void __launch_bounds__(T, (T == 512 ? 2 : (T == 256 ? 3 : 1))) dec_advance_v2_kernel(DecParams p) {
  __shared__ DecShared<T> s;
  B2K_PERSISTENT_LANES((dec_advance_v2_lane<T, PROF, IT>(p, s, lane_, (int)blockIdx.x)))
}

// ------------------------------------------------------------------ finalize (backward sweep)

// link_extra_cost of :342-344 / :433-435
__device__ __forceinline__ float link_extra_cost(float next_extra, float tok_tot, float ac,
                                                 float graph, float next_tot) {
  return next_extra + ((tok_tot + ac + graph) - next_tot);
}

template <int T>
__device__ void dec_finalize_lane(const DecParams &p, const int lane, const int slot) {
  __shared__ uint32_t red32[T / 32];
  __shared__ int redi[T / 32];
  __shared__ int sh_changed;
  const int tid = threadIdx.x;
  const int ch = p.lane_channel[lane];
  ChanState *cs = &p.chan[ch];
  const FstDev &g = p.fst;
  const float kInf = __int_as_float(0x7f800000);
  if (cs->status != B2K_OK || cs->frames_decoded < 0 || cs->finalized) return;

  const int32_t *tok_state = p.tok_state + (size_t)ch * p.max_tokens;
  const float *tok_cost = p.tok_cost + (size_t)ch * p.max_tokens;
  float *tok_extra = p.tok_extra + (size_t)ch * p.max_tokens;
  int4 *links = p.links + (size_t)ch * p.max_links;
  uint32_t *nx = p.new_extra + (size_t)slot * p.max_tpf;
  // A token list that fits (p.fin_scap tokens; one 1024-thread CTA per SM has the room) keeps its per-token sweep state in
  // shared memory: the running minima, their base values, the has-eps flags as a bitmap, and a bitmap of the tokens of
  // list t+1 that survived -- 99.8 % of the links point at dead tokens and now die on a shared-memory bit instead of a
  // random 4-byte read of the arena.  Larger lists use the scratch arrays as before (same code through the pointers).
  extern __shared__ __align__(16) unsigned char fin_smem[];
  const int scap = p.fin_scap, bmw = (scap + 31) / 32;
  uint32_t *nx_s = reinterpret_cast<uint32_t *>(fin_smem), *base_s = nx_s + scap;
  uint32_t *eps_bm = base_s + scap, *alive_next = eps_bm + bmw, *alive_cur = alive_next + bmw;
  bool have_next_bm = false;
  const size_t fo = (size_t)ch * (p.max_frames + 2);
  const int last = cs->frames_decoded;

  // ComputeFinalCosts (:545-586) over the last token list
  const int lb0 = p.frame_tok_begin[fo + last], le0 = p.frame_tok_begin[fo + last + 1];
  uint32_t bc = B2K_INF_ORD, bcf = B2K_INF_ORD;
  int anyf = 0;
  for (int i = lb0 + tid; i < le0; i += T) {
    float fc = __ldg(&g.final_cost[tok_state[i]]);
    float c = tok_cost[i];
    bc = min(bc, f2ord(c));
    bcf = min(bcf, f2ord(c + fc));
    anyf |= (fc != kInf);
  }
  const float best_cost = ord2f(block_min_u32<T>(bc, red32));
  const float best_cost_with_final = ord2f(block_min_u32<T>(bcf, red32));
  const int any_final = block_sum_i32<T>(anyf, redi) > 0;
  const float final_best_cost = (best_cost_with_final != kInf) ? best_cost_with_final : best_cost;

  __shared__ int sh_narcs, sh_nfin;
  if (tid == 0) { sh_narcs = 0; sh_nfin = 0; }
  int n_states = 0;                                    // uniform
  int4 *ls = p.lat_states + (size_t)ch * p.cap_ls;
  int4 *la = p.lat_arcs + (size_t)ch * p.cap_la;
  float2 *lw = p.lat_arcw + (size_t)ch * p.cap_la;
  int2 *lf = p.lat_finals + (size_t)ch * p.cap_lf;
  int *ids_cur = p.cand + (size_t)slot * 5 * p.cand_cap;          // candidate staging is idle here (20*max_tpf ints)
  int *ids_next = ids_cur + p.max_tpf;
  int *has_eps = ids_cur + 2 * p.max_tpf;
  int *surv_e = ids_cur + 3 * p.max_tpf;                         // surviving emitting links of the segment
  const int surv_cap = 5 * p.max_tpf;
  int *surv_p0 = ids_cur + 8 * p.max_tpf, *surv_p1 = ids_cur + 13 * p.max_tpf;   // eps survivors, ping-pong
  __shared__ int sh_cnt[3];
  int tb_next = 0;
  __syncthreads();

  for (int t = last; t >= 0; t--) {
    const int tb = p.frame_tok_begin[fo + t], te = p.frame_tok_begin[fo + t + 1];
    const int n = te - tb;
    if (last > p.max_frames || tb < 0 || te < tb || te > p.max_tokens || n > p.max_tpf) {   // corrupted bookkeeping: fail loudly
      if (tid == 0) {
        printf("b2k dec_finalize: inconsistent frame table ch=%d lane=%d last=%d t=%d tb=%d te=%d ntok=%d\n", ch, lane, last, t, tb, te, cs->ntok);
        { cs->status = B2K_ERR_STATE; cs->err_line = __LINE__; }
      }
      return;
    }
    const int eps_b = p.frame_link_eps[fo + t], eps_e = p.frame_link_begin[fo + t + 1];
    const bool in_s = n <= scap;                              // uniform
    uint32_t *nxp = in_s ? nx_s : nx;
    auto get_eps = [&](int i) -> bool { return in_s ? ((eps_bm[i >> 5] >> (i & 31)) & 1u) != 0u : has_eps[i] != 0; };
    if (in_s) for (int w = tid; w < (n + 31) / 32; w += T) { eps_bm[w] = 0u; alive_cur[w] = 0u; }
    // Almost every token and link dies in this sweep (the lattice keeps ~0.1 % of
    // them), so the passes exit early on dead destinations and keep compact
    // survivor lists instead of flagging or re-reading whole link segments.
    // base value per token: final-cost term on the last list (:426), else
    // +inf (:337) lowered by the emitting links into list t+1
    for (int i = tid; i < n; i += T) {
      float base = kInf;
      if (t == last) {
        float fc = any_final ? __ldg(&g.final_cost[tok_state[tb + i]]) : 0.0f;
        base = tok_cost[tb + i] + fc - final_best_cost;
      }
      nxp[i] = f2ord(base);
      if (!in_s) has_eps[i] = 0;
    }
    if (tid == 0) { sh_cnt[0] = 0; sh_cnt[1] = 0; sh_cnt[2] = 0; }
    __syncthreads();
    if (t < last) {
      const int em_b = p.frame_link_begin[fo + t + 1], em_e = p.frame_link_eps[fo + t + 1];
      constexpr int U = 4;                                    // independent loads in flight per thread
      for (int l0 = em_b; l0 < em_e; l0 += U * T) {
        int4 lk[U];
        float dex[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          int l = l0 + u * T + tid;
          lk[u] = (l < em_e) ? __ldcs(&links[l]) : make_int4(0, -1, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {                         // list t+1 is final
          bool alive = lk[u].y >= 0;
          if (alive && have_next_bm) { const int j = lk[u].y - tb_next; alive = ((alive_next[j >> 5] >> (j & 31)) & 1u) != 0u; }
          dex[u] = alive ? tok_extra[lk[u].y] : kInf;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          if (dex[u] == kInf) continue;                       // link_extra_cost = inf > lattice_beam
          float graph = __int_as_float(__ldg(&g.e_arcs[lk[u].z]).y);
          float lec = link_extra_cost(dex[u], tok_cost[lk[u].x], __int_as_float(lk[u].w), graph, tok_cost[lk[u].y]);
          if (lec > p.lattice_beam) continue;                 // excised (:348-354)
          if (lec < 0.0f) lec = 0.0f;
          atomicMin(&nxp[lk[u].x - tb], f2ord(lec));
          int q = atomicAdd(&sh_cnt[0], 1);
          if (q < surv_cap) surv_e[q] = l0 + u * T + tid;
        }
      }
    }
    for (int l = eps_b + tid; l < eps_e; l += T) {
      const int i = links[l].x - tb;
      if (in_s) atomicOr(&eps_bm[i >> 5], 1u << (i & 31)); else has_eps[i] = 1;
    }
    __syncthreads();
    // tokens without eps out-links already have their exact extra_cost (= base);
    // the others start the eps iteration from the lower bound 0.  The iteration
    // (Jacobi over the surviving eps links; unique fixpoint because eps links
    // form a DAG) only ever excises a link when a LOWER BOUND of its extra cost
    // exceeds lattice_beam, so it keeps exactly the links of :308-379.
    uint32_t *base_ord = in_s ? base_s : reinterpret_cast<uint32_t *>(p.wl + (size_t)slot * 2 * p.max_tpf);
    for (int i = tid; i < n; i += T) {
      uint32_t bo = nxp[i];
      base_ord[i] = bo;
      float v = ord2f(bo);
      if (t == last && v > p.lattice_beam) v = kInf;
      tok_extra[tb + i] = get_eps(i) ? 0.0f : v;
    }
    __syncthreads();
    int n_eps_in = eps_e - eps_b;                             // iteration 0 reads the segment itself
    int *eps_in = nullptr, *eps_out = surv_p0;
    for (int iter = 0; iter < 100000 && n_eps_in > 0; iter++) {
      if (tid == 0) { sh_changed = 0; sh_cnt[1] = 0; }
      __syncthreads();
      for (int k = tid; k < n_eps_in; k += T) {
        int l = eps_in ? eps_in[k] : eps_b + k;
        int4 lk = links[l];
        float dex = tok_extra[lk.y];
        if (dex == kInf) continue;
        float graph = __int_as_float(__ldg(&g.ne_arcs[(uint32_t)lk.z & B2K_ARC_MASK]).y);
        float lec = link_extra_cost(dex, tok_cost[lk.x], 0.0f, graph, tok_cost[lk.y]);
        if (lec > p.lattice_beam) continue;
        if (lec < 0.0f) lec = 0.0f;
        atomicMin(&nxp[lk.x - tb], f2ord(lec));
        int q = atomicAdd(&sh_cnt[1], 1);
        if (q < surv_cap) eps_out[q] = l;
      }
      __syncthreads();
      const int n_out = min(sh_cnt[1], surv_cap);
      int changed = 0;
      // only sources of eps links can change: re-evaluate them via the survivor list plus the
      // sources whose links all died (their value falls back to base) -> walk the has_eps tokens
      for (int i = tid; i < n; i += T) {
        if (!get_eps(i)) continue;
        float v = ord2f(nxp[i]);
        if (t == last && v > p.lattice_beam) v = kInf;      // :458-459
        float old = tok_extra[tb + i];
        if (!(v == old)) changed = 1;
        tok_extra[tb + i] = v;
        nxp[i] = base_ord[i];
      }
      if (changed) sh_changed = 1;
      __syncthreads();
      const int ch_any = sh_changed;
      if (tid == 0 && sh_cnt[1] > surv_cap) { cs->status = B2K_ERR_OVERFLOW; cs->err_line = __LINE__; }
      eps_in = eps_out;
      eps_out = (eps_out == surv_p0) ? surv_p1 : surv_p0;
      n_eps_in = n_out;
      __syncthreads();
      if (!ch_any) break;
    }
    // ---- lattice-state ids for the surviving tokens of list t (ids grow as we walk
    //      backwards; the host flips them so that id 0 is the start state) and emission of
    //      the surviving states / arcs into the channel's compact lattice region
    // The ids are flipped when the lattice is packed (the last id becomes state 0) and state 0 must be the START state
    // (include/b2k.h; the determinizer starts there): on the first token list the token of the graph's start state takes its
    // id after every other survivor.  (Until round 2 it took whatever the atomic handed it: whenever frame 0 kept an
    // eps-successor of the start as well, state 0 was one of the two at random -- found as a flaky difference between two
    // determinizations of the same utterance, tools/determinism_probe.py; the order-free comparisons never saw it.)
    __shared__ int sh_start_k;
    if (t == 0 && tid == 0) sh_start_k = -1;
    if (t == 0) __syncthreads();
    for (int k = tid; k < n; k += T) {
      float ex = tok_extra[tb + k];
      if (ex == kInf) continue;
      if (t == 0 && tok_state[tb + k] == g.start) { sh_start_k = k; continue; }
      int id = n_states + atomicAdd(&sh_cnt[2], 1);
      ids_cur[k] = id;
      if (in_s) atomicOr(&alive_cur[k >> 5], 1u << (k & 31));
      if (id < p.cap_ls) ls[id] = make_int4(t, tok_state[tb + k], __float_as_int(tok_cost[tb + k]), __float_as_int(ex));
      if (t == last) {
        float fc = any_final ? __ldg(&g.final_cost[tok_state[tb + k]]) : 0.0f;
        if (fc != kInf) {
          int q = atomicAdd(&sh_nfin, 1);
          if (q < p.cap_lf) lf[q] = make_int2(id, __float_as_int(fc));
        }
      }
    }
    __syncthreads();
    if (t == 0 && sh_start_k >= 0) {                          // the start state's token: the last id of the lattice
      if (tid == 0) {
        const int k = sh_start_k;
        const int id = n_states + sh_cnt[2];
        sh_cnt[2] += 1;
        ids_cur[k] = id;
        if (in_s) atomicOr(&alive_cur[k >> 5], 1u << (k & 31));
        if (id < p.cap_ls) ls[id] = make_int4(t, tok_state[tb + k], __float_as_int(tok_cost[tb + k]), __float_as_int(tok_extra[tb + k]));
        if (t == last) {                                      // (a zero-frame utterance: the first list is the last one too)
          const float fc = any_final ? __ldg(&g.final_cost[tok_state[tb + k]]) : 0.0f;
          if (fc != kInf) {
            const int q = atomicAdd(&sh_nfin, 1);
            if (q < p.cap_lf) lf[q] = make_int2(id, __float_as_int(fc));
          }
        }
      }
      __syncthreads();
    }
    n_states += sh_cnt[2];
    if (t < last) {
      const float coff = p.frame_cost_offset[(size_t)ch * (p.max_frames + 1) + t];
      const int ns = min(sh_cnt[0], surv_cap);
      if (tid == 0 && sh_cnt[0] > surv_cap) { cs->status = B2K_ERR_OVERFLOW; cs->err_line = __LINE__; }
      for (int k = tid; k < ns; k += T) {
        int4 lk = links[surv_e[k]];
        int4 arc = __ldg(&g.e_arcs[lk.z]);
        int q = atomicAdd(&sh_narcs, 1);
        if (q < p.cap_la) {
          la[q] = make_int4(ids_cur[lk.x - tb], ids_next[lk.y - tb_next], __ldg(&g.e_ilabel[lk.z]), arc.w);
          lw[q] = make_float2(__int_as_float(arc.y), __int_as_float(lk.w) - coff);   // :174-181
        }
      }
    }
    // surviving eps links: the last input list of the iteration, re-checked against the final extras
    for (int k = tid; k < n_eps_in; k += T) {
      int l = eps_in ? eps_in[k] : eps_b + k;
      int4 lk = links[l];
      float dex = tok_extra[lk.y];
      if (dex == kInf) continue;
      int4 arc = __ldg(&g.ne_arcs[(uint32_t)lk.z & B2K_ARC_MASK]);
      float lec = link_extra_cost(dex, tok_cost[lk.x], 0.0f, __int_as_float(arc.y), tok_cost[lk.y]);
      if (lec > p.lattice_beam) continue;
      int q = atomicAdd(&sh_narcs, 1);
      if (q < p.cap_la) {
        la[q] = make_int4(ids_cur[lk.x - tb], ids_cur[lk.y - tb], 0, arc.z & 0x7fffffff);
        lw[q] = make_float2(__int_as_float(arc.y), 0.0f);
      }
    }
    __syncthreads();
    { int *tmp = ids_cur; ids_cur = ids_next; ids_next = tmp; }
    { uint32_t *tmp = alive_cur; alive_cur = alive_next; alive_next = tmp; }
    have_next_bm = in_s;
    tb_next = tb;
  }
  if (tid == 0) {
    cs->finalized = 1;
    cs->lat_states = n_states;
    cs->lat_arcs = sh_narcs;
    cs->lat_finals = sh_nfin;
    cs->any_final = any_final;
    cs->final_best_cost = final_best_cost;
    if (n_states > p.cap_ls || sh_narcs > p.cap_la || sh_nfin > p.cap_lf) { cs->status = B2K_ERR_OVERFLOW; cs->err_line = __LINE__; }
  }
}

template <int T>
__global__ void __launch_bounds__(T) dec_finalize_kernel(DecParams p) {
  B2K_PERSISTENT_LANES(dec_finalize_lane<T>(p, lane_, (int)blockIdx.x))
}

// ------------------------------------------------------------------ lattice packing

// Copies the compact lattices of n channels into packed arrays (one D2H each).
// offs: [3*(n+1)] exclusive prefix sums of states / arcs / finals, computed on the host.
__global__ void dec_pack_kernel(DecParams p, const int32_t *channels, const int64_t *offs, int n,
                                int4 *o_states, int4 *o_arcs, float2 *o_arcw, int2 *o_finals) {
  const int c = blockIdx.y;
  const int ch = channels[c];
  const ChanState *cs = &p.chan[ch];
  const int ns = min(cs->lat_states, p.cap_ls), na = min(cs->lat_arcs, p.cap_la), nf = min(cs->lat_finals, p.cap_lf);
  const int4 *ls = p.lat_states + (size_t)ch * p.cap_ls;
  const int4 *la = p.lat_arcs + (size_t)ch * p.cap_la;
  const float2 *lw = p.lat_arcw + (size_t)ch * p.cap_la;
  const int2 *lf = p.lat_finals + (size_t)ch * p.cap_lf;
  const int64_t so = offs[c], ao = offs[(n + 1) + c], fo = offs[2 * (n + 1) + c];
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
  // flip ids so that lattice state 0 is the start state (ids were assigned walking backwards)
  for (int i = gt; i < ns; i += gs) o_states[so + (ns - 1 - i)] = ls[i];
  for (int i = gt; i < na; i += gs) {
    int4 a = la[i];
    a.x = ns - 1 - a.x; a.y = ns - 1 - a.y;
    o_arcs[ao + i] = a; o_arcw[ao + i] = lw[i];
  }
  for (int i = gt; i < nf; i += gs) { int2 f = lf[i]; f.x = ns - 1 - f.x; o_finals[fo + i] = f; }
}

__global__ void dec_reset_channels_kernel(DecParams p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ChanState *cs = &p.chan[p.lane_channel[i]];
  ChanState z;
  memset(&z, 0, sizeof(z));
  z.frames_decoded = -1;
  *cs = z;
}

// ---- the same packing without a host round trip: offsets computed on the device into the buffer's header.
// Buffer layout: int64 header[3 * (n + 1) + 4] = {state offsets, arc offsets, final offsets, status, bytes needed, n, first failing entry (line << 32 | index) or -1},
// padded to 16 bytes, then int4 states[ns], int4 arcs[na], float2 arc weights[na], int2 finals[nf].
__host__ __device__ inline int64_t pack_header_bytes(int n) { return (int64_t)((8 * (3 * ((int64_t)n + 1) + 4) + 15) / 16 * 16); }

__global__ void dec_pack_header_kernel(DecParams p, const int32_t *channels, int n, int64_t *hdr, int64_t cap_bytes) {
  // one block: exclusive prefix sums of the per-channel lattice sizes (n is at most a few thousand)
  __shared__ long long carry[3];
  __shared__ int status;
  __shared__ long long culprit;                               // first failing entry: (source line << 32) | index in `channels`
  if (threadIdx.x == 0) { carry[0] = carry[1] = carry[2] = 0; status = B2K_OK; culprit = -1; }
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    int v[3] = {0, 0, 0};
    if (i < n) {
      const ChanState *cs = &p.chan[channels[i]];
      if (cs->status != B2K_OK) { if (atomicCAS(&status, B2K_OK, cs->status) == B2K_OK) culprit = ((long long)cs->err_line << 32) | (unsigned)i; }
      else if (!cs->finalized) { if (atomicCAS(&status, B2K_OK, B2K_ERR_STATE) == B2K_OK) culprit = (long long)(unsigned)i; }
      v[0] = min(cs->lat_states, p.cap_ls); v[1] = min(cs->lat_arcs, p.cap_la); v[2] = min(cs->lat_finals, p.cap_lf);
    }
    for (int k = 0; k < 3; k++) {
      // block-wide exclusive scan by a shared array (n is small; simplicity over speed)
      __shared__ int sc[1024];
      sc[threadIdx.x] = v[k];
      __syncthreads();
      for (int o = 1; o < (int)blockDim.x; o <<= 1) {
        int t = threadIdx.x >= (unsigned)o ? sc[threadIdx.x - o] : 0;
        __syncthreads();
        sc[threadIdx.x] += t;
        __syncthreads();
      }
      if (i < n) hdr[(size_t)k * (n + 1) + i] = carry[k] + sc[threadIdx.x] - v[k];
      __syncthreads();
      if (threadIdx.x == blockDim.x - 1) carry[k] += sc[threadIdx.x];
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    for (int k = 0; k < 3; k++) hdr[(size_t)k * (n + 1) + n] = carry[k];
    const int64_t need = pack_header_bytes(n) + carry[0] * 16 + carry[1] * 24 + carry[2] * 8;
    if (need > cap_bytes && status == B2K_OK) status = B2K_ERR_OVERFLOW;
    int64_t *tail = hdr + 3 * ((size_t)n + 1);
    tail[0] = status; tail[1] = need; tail[2] = n; tail[3] = culprit;
  }
}

__global__ void dec_pack_body_kernel(DecParams p, const int32_t *channels, int n, const int64_t *hdr, char *buf) {
  if (hdr[3 * ((size_t)n + 1)] != B2K_OK) return;             // error or overflow: the header says so, nothing is written
  const int64_t ns_all = hdr[n], na_all = hdr[(n + 1) + n];
  int4 *o_states = reinterpret_cast<int4 *>(buf + pack_header_bytes(n));
  int4 *o_arcs = o_states + ns_all;
  float2 *o_arcw = reinterpret_cast<float2 *>(o_arcs + na_all);
  int2 *o_finals = reinterpret_cast<int2 *>(o_arcw + na_all);
  const int c = blockIdx.y;
  const int ch = channels[c];
  const ChanState *cs = &p.chan[ch];
  const int ns = min(cs->lat_states, p.cap_ls), na = min(cs->lat_arcs, p.cap_la), nf = min(cs->lat_finals, p.cap_lf);
  const int4 *ls = p.lat_states + (size_t)ch * p.cap_ls;
  const int4 *la = p.lat_arcs + (size_t)ch * p.cap_la;
  const float2 *lw = p.lat_arcw + (size_t)ch * p.cap_la;
  const int2 *lf = p.lat_finals + (size_t)ch * p.cap_lf;
  const int64_t so = hdr[c], ao = hdr[(n + 1) + c], fo = hdr[2 * (n + 1) + c];
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
  for (int i = gt; i < ns; i += gs) o_states[so + (ns - 1 - i)] = ls[i];
  for (int i = gt; i < na; i += gs) {
    int4 a = la[i];
    a.x = ns - 1 - a.x; a.y = ns - 1 - a.y;
    o_arcs[ao + i] = a; o_arcw[ao + i] = lw[i];
  }
  for (int i = gt; i < nf; i += gs) { int2 f = lf[i]; f.x = ns - 1 - f.x; o_finals[fo + i] = f; }
}


// ---------------------------------------------------------------------------------------------------------------------
// Best path of a channel WITHOUT finalizing it: LatticeFasterOnlineDecoderTpl::BestPathEnd + TraceBackBestPath
// (decoder/lattice-faster-online-decoder.cc:78-167), ComputeFinalCosts' final_relative_cost (lattice-faster-decoder.cc:
// 545-586) -- what partial hypotheses, endpointing (online2/online-endpoint.cc:78-110: trailing silence of the best path
// with use_final_probs = false) and CudaDecoder::GetBestPath (cudadecoder/cuda-decoder.h:279) need mid-utterance.
// The tokens here carry no back pointer (the reference's non-online Token does not either): the predecessor of a token is
// the link into it with the smallest (source cost + link cost) -- the link that set the token's cost; ties go to the lowest
// arena index.  One CTA per channel walks back from the best token of the last list; every step scans the links of one
// frame (coalesced 16-byte records) and reduces a 64-bit (cost, index) key.
struct BestPathHdr {
  int32_t status, n_arcs, end_state, frames;
  float final_cost, best_cost, final_relative_cost, pad;
};

template <int T>
__global__ void __launch_bounds__(T) dec_best_path_kernel(DecParams p, const int32_t *channels, int use_final, int cap,
                                                          int4 *out_arcs, int2 *out_where, BestPathHdr *out_hdr) {
  __shared__ unsigned long long sh64[T / 32];
  const int tid = threadIdx.x;
  const int ch = channels[blockIdx.x];
  const FstDev &g = p.fst;
  const ChanState *cs = &p.chan[ch];
  const float kInf = __int_as_float(0x7f800000);
  int4 *oa = out_arcs + (size_t)blockIdx.x * cap;
  int2 *ow = out_where + (size_t)blockIdx.x * cap;
  BestPathHdr h;
  h.status = 0; h.n_arcs = 0; h.end_state = -1; h.frames = cs->frames_decoded;
  h.final_cost = 0.0f; h.best_cost = kInf; h.final_relative_cost = kInf; h.pad = 0.0f;
  const int last = cs->frames_decoded;
  if (last < 0 || last > p.max_frames || cs->status != 0) {
    h.status = B2K_ERR_STATE;
    if (tid == 0) out_hdr[blockIdx.x] = h;
    return;
  }
  const int32_t *tok_state = p.tok_state + (size_t)ch * p.max_tokens;
  const float *tok_cost = p.tok_cost + (size_t)ch * p.max_tokens;
  const int4 *links = p.links + (size_t)ch * p.max_links;
  const size_t fo = (size_t)ch * (p.max_frames + 2);
  const float *coffs = p.frame_cost_offset + (size_t)ch * (p.max_frames + 1);
  // the end token (BestPathEnd :78-117) and final_relative_cost
  const int tb = p.frame_tok_begin[fo + last], te = p.frame_tok_begin[fo + last + 1];
  unsigned long long ka = ~0ull, kb = ~0ull;
  for (int i = tb + tid; i < te; i += T) {
    const float c = tok_cost[i], fc = __ldg(&g.final_cost[tok_state[i]]);
    const unsigned long long a = ((unsigned long long)f2ord(c) << 32) | (uint32_t)(i - tb);
    const unsigned long long b = ((unsigned long long)f2ord(c + fc) << 32) | (uint32_t)(i - tb);
    ka = a < ka ? a : ka; kb = b < kb ? b : kb;
  }
  ka = block_min_u64<T>(ka, sh64);
  kb = block_min_u64<T>(kb, sh64);
  const float best_cost = (ka == ~0ull) ? kInf : ord2f((uint32_t)(ka >> 32));
  const float best_with_final = (kb == ~0ull) ? kInf : ord2f((uint32_t)(kb >> 32));
  h.final_relative_cost = (best_cost == kInf && best_with_final == kInf) ? kInf : best_with_final - best_cost;
  const bool with_final = use_final && best_with_final != kInf;     // "any final tokens were active on the final frame"
  const unsigned long long kend = with_final ? kb : ka;
  h.best_cost = with_final ? best_with_final : best_cost;
  if (kend == ~0ull || h.best_cost == kInf) {                       // "No final token found."
    if (tid == 0) out_hdr[blockIdx.x] = h;
    return;
  }
  int cur = tb + (int)(uint32_t)(kend & 0xffffffffull);
  h.end_state = tok_state[cur];
  h.final_cost = with_final ? __ldg(&g.final_cost[h.end_state]) : 0.0f;
  int t = last, n = 0;
  for (;;) {
    const int lb = p.frame_link_begin[fo + t], lm = p.frame_link_eps[fo + t], le = p.frame_link_begin[fo + t + 1];
    unsigned long long best = ~0ull;
    for (int l = lb + tid; l < le; l += T) {
      const int4 lk = links[l];
      if (lk.y != cur) continue;
      float v;
      if (l < lm) v = (tok_cost[lk.x] + __int_as_float(lk.w)) + __int_as_float(__ldg(&g.e_arcs[lk.z]).y);
      else v = tok_cost[lk.x] + __int_as_float(__ldg(&g.ne_arcs[(uint32_t)lk.z & B2K_ARC_MASK]).y);
      const unsigned long long k = ((unsigned long long)f2ord(v) << 32) | (uint32_t)(l - lb);
      best = k < best ? k : best;
    }
    best = block_min_u64<T>(best, sh64);
    const bool found = best != ~0ull;
    if (t == 0 && (!found || ord2f((uint32_t)(best >> 32)) > tok_cost[cur])) break;    // the start token: no link made it
    if (!found) { h.status = B2K_ERR_STATE; break; }
    if (n >= cap) { h.status = B2K_ERR_OVERFLOW; break; }
    const int l = lb + (int)(uint32_t)(best & 0xffffffffull);
    const int4 lk = links[l];
    if (tid == 0) {
      if (l < lm) {
        const int4 arc = __ldg(&g.e_arcs[lk.z]);
        oa[n] = make_int4(__ldg(&g.e_ilabel[lk.z]), arc.w, arc.y, __float_as_int(__int_as_float(lk.w) - coffs[t - 1]));   // :150-154
      } else {
        const int4 arc = __ldg(&g.ne_arcs[(uint32_t)lk.z & B2K_ARC_MASK]);
        oa[n] = make_int4(0, arc.z & 0x7fffffff, arc.y, 0);
      }
      ow[n] = make_int2(t, tok_state[cur]);
    }
    n++;
    cur = lk.x;
    if (l < lm) t--;
  }
  if (h.status == 0 && tok_state[cur] != g.start) h.status = B2K_ERR_STATE;           // the walk must end in the start state
  h.n_arcs = n;
  if (tid == 0) out_hdr[blockIdx.x] = h;
}

}  // namespace b2k

// ====================================================================== host side

using namespace b2k;

struct b2k_fst {
  FstDev dev;
  int2 *d_st_off = nullptr;
  int4 *d_e = nullptr, *d_ne = nullptr;
  int32_t *d_eil = nullptr;
  float *d_final = nullptr;
  // host copies for debug/extraction
  std::vector<int4> h_e, h_ne;
  std::vector<int32_t> h_eil;
  int32_t num_pdfs_seen = 0;
};

struct b2k_dec {
  const b2k_fst *fst;
  b2k_dec_cfg cfg;
  int nlanes, nchannels;
  DecParams p;
  int threads_override = 0, fin_threads = 1024, num_sms = 0;   // tuning knobs, read from the environment at creation
  int ctas_override = 0;                // B2K_DEC_CTAS (256-thread CTAs only)
  bool cid_smem_off = false;            // B2K_DEC_CID_SMEM=0
  bool par_walk_off = true;             // B2K_DEC_PARWALK=1: replay by connected components (bit-exact, measured no faster: one component holds most of a heavy frame's tokens)
  bool fin_smem_off = true;             // B2K_FIN_SMEM=1 turns the shared-memory sweep state on (measured slower: 94 vs 58 ms, it shrinks L1 to 92 KB)
  int arcs_per_thread = 3;              // B2K_DEC_IT (512-thread CTAs): 3 measured 1.7 % faster than 4 and 2 (r2n, r2o)
  int nslots = 0;                       // scratch slots = the largest grid any per-lane launch uses (resident CTAs)
  int prof = 0;                         // B2K_DEC_PROF=1: per-phase cycle counters in the reference-order kernel
  int ll_smem_off = 0;                  // B2K_DEC_LL_SMEM=0: leave the log-likelihood rows in global memory
  int use_v2 = 0;                       // second-generation reference-order frame step (B2K_DEC_V1=1 selects the first)
  int grid_cap = 0;                     // B2K_DEC_GRID: cap on the resident CTAs of the reference-order launches (experiments)
  int32_t *d_lane_counter = nullptr;
  // launch-argument staging
  int32_t *d_lane_channel = nullptr;
  const float **d_lane_ll = nullptr;
  int32_t *d_lane_nframes = nullptr;
  int32_t *h_lane_channel = nullptr;    // pinned
  const float **h_lane_ll = nullptr;
  int32_t *h_lane_nframes = nullptr;
  cudaEvent_t staging_free = nullptr;
  std::vector<void *> allocs;
  // batched lattice read-back (grow-only scratch)
  char *d_pack = nullptr; size_t d_pack_bytes = 0;
  char *h_pack = nullptr; size_t h_pack_bytes = 0;
  int32_t *d_pack_ch = nullptr; int64_t *d_pack_offs = nullptr;
  ChanState *h_chan = nullptr;
  // best-path read-back (grow-only scratch)
  char *d_bp = nullptr; size_t d_bp_bytes = 0;
  char *h_bp = nullptr; size_t h_bp_bytes = 0;
};

extern "C" {

void b2k_dec_cfg_default(b2k_dec_cfg *c) {
  if (!c) return;
  c->beam = 15.0f; c->lattice_beam = 8.0f; c->max_active = 7000; c->min_active = 200;
  c->beam_delta = 0.5f; c->prune_interval = 25; c->prune_scale = 0.1f;
  c->max_tokens_per_frame = 32768; c->max_frames = 1024;
  c->max_tokens = 3000000; c->max_links = 6000000;
  c->reference_order = 1; c->hash_ratio = 2.0f; c->max_arcs_per_frame = 1 << 20;
  c->max_lattice_states = 131072; c->max_lattice_arcs = 262144;
}

int b2k_fst_create(const b2k_fst_csr *csr, b2k_fst **out) {
  if (!csr || !out || csr->num_states <= 0) return set_error(B2K_ERR_INVALID, "b2k_fst_create: bad args");
  {   // the CSR comes from the caller: everything the kernels will index with is checked here, before any upload
    const int N0 = csr->num_states;
    if (!csr->offsets || !csr->final_cost || csr->start < 0 || csr->start >= N0 || csr->offsets[0] != 0)
      return set_error(B2K_ERR_INVALID, "b2k_fst_create: bad start state or offsets");
    for (int s = 0; s < N0; s++) if (csr->offsets[s + 1] < csr->offsets[s]) return set_error(B2K_ERR_INVALID, "b2k_fst_create: arc offsets are not non-decreasing");
    const int A0 = csr->offsets[N0];
    if (A0 > 0 && (!csr->ilabel || !csr->olabel || !csr->weight || !csr->nextstate)) return set_error(B2K_ERR_INVALID, "b2k_fst_create: arc arrays missing");
    if (csr->tid2pdf && csr->num_tids <= 0) return set_error(B2K_ERR_INVALID, "b2k_fst_create: empty transition-id table");
    for (int a = 0; a < A0; a++) {
      if (csr->nextstate[a] < 0 || csr->nextstate[a] >= N0) return set_error(B2K_ERR_INVALID, "b2k_fst_create: arc to a state outside the graph");
      if (csr->ilabel[a] < 0 || csr->olabel[a] < 0) return set_error(B2K_ERR_INVALID, "b2k_fst_create: negative label");
      if (csr->weight[a] != csr->weight[a]) return set_error(B2K_ERR_INVALID, "b2k_fst_create: NaN arc weight");
    }
  }
  int rc = require_device();
  if (rc) return rc;
  const int N = csr->num_states;
  b2k_fst *f = new b2k_fst();
  std::vector<int2> st_off(N + 1);
  int ne_cnt = 0, e_cnt = 0;
  for (int s = 0; s < N; s++) {
    st_off[s] = make_int2(e_cnt, ne_cnt);
    for (int a = csr->offsets[s]; a < csr->offsets[s + 1]; a++) {
      if (csr->ilabel[a] == 0) ne_cnt++; else e_cnt++;
    }
  }
  st_off[N] = make_int2(e_cnt, ne_cnt);
  if ((uint32_t)e_cnt > B2K_ARC_MASK || (uint32_t)ne_cnt > B2K_ARC_MASK) { delete f; return set_error(B2K_ERR_INVALID, "graphs above 2^30 arcs per class are not supported"); }
  f->h_e.resize(e_cnt); f->h_ne.resize(ne_cnt); f->h_eil.resize(e_cnt);
  int ei = 0, ni = 0;
  for (int s = 0; s < N; s++) {
    for (int a = csr->offsets[s]; a < csr->offsets[s + 1]; a++) {
      int wbits; memcpy(&wbits, &csr->weight[a], 4);
      int il = csr->ilabel[a];
      if (il == 0) {
        f->h_ne[ni++] = make_int4(csr->nextstate[a], wbits, csr->olabel[a], 0);
      } else {
        if (csr->tid2pdf && (il < 0 || il >= csr->num_tids)) { delete f; return set_error(B2K_ERR_INVALID, "ilabel out of tid2pdf range"); }
        int pdf = csr->tid2pdf ? csr->tid2pdf[il] : il - 1;   // cuda-fst.cc:166-175
        f->h_eil[ei] = il;
        f->h_e[ei++] = make_int4(csr->nextstate[a], wbits, pdf, csr->olabel[a]);
        if (pdf + 1 > f->num_pdfs_seen) f->num_pdfs_seen = pdf + 1;
      }
    }
  }
  // eps arcs: flag the last eps arc of each state and store where the destination's eps arcs start
  // (lets the sequential eps replay run without touching the state table)
  for (int s = 0; s < N; s++) {
    int b = st_off[s].y, e = st_off[s + 1].y;
    for (int a = b; a < e; a++) {
      int ns = f->h_ne[a].x;
      int nb = st_off[ns].y, ne = st_off[ns + 1].y;
      f->h_ne[a].w = (ne > nb) ? nb : -1;
      if (a == e - 1) f->h_ne[a].z = (int)((uint32_t)f->h_ne[a].z | 0x80000000u);
    }
  }
  auto up = [&](void **d, const void *h, size_t bytes) -> int {
    B2K_CUDA_CHECK(cudaMalloc(d, bytes ? bytes : 16));
    if (bytes) B2K_CUDA_CHECK(cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice));
    return 0;
  };
  if ((rc = up((void **)&f->d_st_off, st_off.data(), sizeof(int2) * (N + 1)))) return rc;
  if ((rc = up((void **)&f->d_e, f->h_e.data(), sizeof(int4) * e_cnt))) return rc;
  if ((rc = up((void **)&f->d_ne, f->h_ne.data(), sizeof(int4) * ne_cnt))) return rc;
  if ((rc = up((void **)&f->d_eil, f->h_eil.data(), 4 * (size_t)e_cnt))) return rc;
  if ((rc = up((void **)&f->d_final, csr->final_cost, 4 * (size_t)N))) return rc;
  f->dev.num_states = N; f->dev.start = csr->start; f->dev.num_e = e_cnt; f->dev.num_ne = ne_cnt;
  f->dev.st_off = f->d_st_off; f->dev.e_arcs = f->d_e; f->dev.ne_arcs = f->d_ne;
  f->dev.e_ilabel = f->d_eil; f->dev.final_cost = f->d_final;
  *out = f;
  return B2K_OK;
}

int b2k_fst_destroy(b2k_fst *f) {
  if (!f) return B2K_OK;
  cudaFree(f->d_st_off); cudaFree(f->d_e); cudaFree(f->d_ne); cudaFree(f->d_eil); cudaFree(f->d_final);
  delete f;
  return B2K_OK;
}
int32_t b2k_fst_num_states(const b2k_fst *f) { return f ? f->dev.num_states : -1; }
int32_t b2k_fst_start(const b2k_fst *f) { return f ? f->dev.start : -1; }

static int ilog2_ceil(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

static int dec_create_impl(b2k_dec *d, const b2k_fst *fst, const b2k_dec_cfg *cfg, int32_t nlanes, int32_t nchannels);

int b2k_dec_create(const b2k_fst *fst, const b2k_dec_cfg *cfg, int32_t nlanes, int32_t nchannels,
                   b2k_dec **out) {
  if (!fst || !cfg || !out || nlanes <= 0 || nchannels < nlanes)
    return set_error(B2K_ERR_INVALID, "b2k_dec_create: bad args");
  if (!(cfg->beam > 0.0f && cfg->max_active > 1 && cfg->lattice_beam > 0.0f &&
        cfg->min_active <= cfg->max_active && cfg->beam_delta > 0.0f &&
        (!cfg->reference_order || cfg->hash_ratio >= 1.0f)))   // Check() lattice-faster-decoder.h:99-105
    return set_error(B2K_ERR_INVALID, "b2k_dec_create: invalid decoder config");
  if (cfg->max_tokens > 0x7fffffffLL || cfg->max_links > 0x7fffffffLL)
    return set_error(B2K_ERR_INVALID, "arena capacities must fit int32");
  int rc = require_device();
  if (rc) return rc;
  b2k_dec *d = new b2k_dec();
  rc = dec_create_impl(d, fst, cfg, nlanes, nchannels);
  if (rc) {                       // e.g. cudaMalloc failed half way (arenas are sized by the caller): release what exists
    b2k_dec_destroy(d);
    *out = nullptr;
    return rc;
  }
  *out = d;
  return B2K_OK;
}

static int dec_create_impl(b2k_dec *d, const b2k_fst *fst, const b2k_dec_cfg *cfg, int32_t nlanes, int32_t nchannels) {
  int rc = 0;
  d->fst = fst; d->cfg = *cfg; d->nlanes = nlanes; d->nchannels = nchannels;
  {
    // tuning knobs (defaults = measured best); read per decoder so that tests can vary them
    if (const char *e = getenv("B2K_DEC_THREADS")) { int v = atoi(e); if (v == 128 || v == 256 || v == 512 || v == 1024) d->threads_override = v; }
    d->fin_threads = 1024;   // one 1024-thread CTA per SM: measured 4x faster than four 256-thread CTAs
    if (const char *e = getenv("B2K_FIN_THREADS")) { int v = atoi(e); if (v == 256 || v == 512 || v == 1024) d->fin_threads = v; }
    int dev = 0;
    B2K_CUDA_CHECK(cudaGetDevice(&dev));
    B2K_CUDA_CHECK(cudaDeviceGetAttribute(&d->num_sms, cudaDevAttrMultiProcessorCount, dev));
    if (const char *e = getenv("B2K_DEC_PROF")) d->prof = atoi(e) != 0;
    if (const char *e = getenv("B2K_DEC_GRID")) d->grid_cap = atoi(e);
    if (const char *e = getenv("B2K_DEC_LL_SMEM")) d->ll_smem_off = atoi(e) == 0;
    if (const char *e = getenv("B2K_DEC_CID_SMEM")) d->cid_smem_off = atoi(e) == 0;
    if (const char *e = getenv("B2K_DEC_PARWALK")) d->par_walk_off = atoi(e) == 0;
    if (const char *e = getenv("B2K_FIN_SMEM")) d->fin_smem_off = atoi(e) == 0;
    if (const char *e = getenv("B2K_DEC_IT")) { int v = atoi(e); if (v >= 2 && v <= 4) d->arcs_per_thread = v; }
    // Persistent launches: at most two CTAs per SM are ever resident (512-thread reference-order CTAs; the
    // order-free and finalize kernels use one slot per CTA of their own, smaller grids), so that is the
    // number of scratch slots, whatever the batch size.
    if (const char *e = getenv("B2K_DEC_CTAS")) { int v = atoi(e); if (v >= 1 && v <= 3) d->ctas_override = v; }   // experiments: 256-thread CTAs, three per SM
    d->nslots = std::min(nlanes, std::max(2, d->ctas_override) * d->num_sms);
  }

  DecParams &p = d->p;
  memset(&p, 0, sizeof(p));
  p.fst = fst->dev;
  p.beam = cfg->beam; p.lattice_beam = cfg->lattice_beam; p.beam_delta = cfg->beam_delta;
  p.max_active = cfg->max_active; p.min_active = cfg->min_active;
  p.hash_log = ilog2_ceil(std::max(cfg->max_tokens_per_frame, 1024) * 2);
  p.hash_size = 1 << p.hash_log;
  p.max_tpf = p.hash_size / 2;
  p.cand_cap = p.max_tpf * 4;
  p.max_frames = cfg->max_frames;
  p.max_tokens = (int32_t)cfg->max_tokens; p.max_links = (int32_t)cfg->max_links;
  auto alloc = [&](void **ptr, size_t bytes, int fill) -> int {
    B2K_CUDA_CHECK(cudaMalloc(ptr, bytes));
    d->allocs.push_back(*ptr);
    B2K_CUDA_CHECK(cudaMemset(*ptr, fill, bytes));
    return 0;
  };
  size_t nc = nchannels, nl = (size_t)d->nslots, nlanes_sz = (size_t)nlanes;
#define A(ptr, bytes, fill) if ((rc = alloc((void **)&(ptr), (bytes), (fill)))) return rc;
  A(p.chan, sizeof(ChanState) * nc, 0);
  A(p.tok_state, 4 * nc * p.max_tokens, 0);
  A(p.tok_cost, 4 * nc * p.max_tokens, 0);
  A(p.tok_extra, 4 * nc * p.max_tokens, 0);
  A(p.links, sizeof(int4) * nc * p.max_links, 0);
  A(p.frame_tok_begin, 4 * nc * (p.max_frames + 2), 0);
  A(p.frame_link_begin, 4 * nc * (p.max_frames + 2), 0);
  A(p.frame_link_eps, 4 * nc * (p.max_frames + 2), 0);
  A(p.frame_cost_offset, 4 * nc * (p.max_frames + 1), 0);
  A(p.frame_cutoff, 4 * nc * (p.max_frames + 1), 0);
  A(p.hash, sizeof(int4) * nl * p.hash_size, 0);
  A(p.tokslot, 4 * nl * p.max_tpf, 0);
  A(p.wl, 4 * nl * 2 * p.max_tpf, 0);
  A(p.cand, 4 * nl * 5 * (size_t)p.cand_cap, 0);
  A(p.new_extra, 4 * nl * p.max_tpf, 0);
  A(p.lane_stamp, 4 * nl, 0);
  if (cfg->reference_order) {
    p.pos_cap = ((std::max(cfg->max_arcs_per_frame, 1024) + 31) / 32) * 32;
    p.hc_cap = std::max(1000, (int)((float)p.max_tpf * cfg->hash_ratio) + 1);
    p.queue_cap = 5 * p.cand_cap;
    p.hash_ratio = cfg->hash_ratio;
    A(p.x_bm, 4 * nl * (p.pos_cap / 32), 0);
    A(p.x_wbase, 4 * nl * (p.pos_cap / 32), 0);
    A(p.x_by_ins, 4 * nl * p.max_tpf, 0);
    A(p.x_bk, sizeof(int4) * nl * p.hc_cap, 0);
    A(p.x_sbase, 4 * nl * p.max_tpf, 0);
    A(p.x_run, 4 * nl * p.max_tpf, 0);
    A(p.x_order, 4 * nl * p.max_tpf, 0);
    p.adj_cap = 2 * p.max_tpf;
    p.rs_rcap = 4096; p.rs_ecap = 4096; p.rs_qcap = 4096;  // shared-memory walk: 80 KB per CTA (two 512-thread CTAs per SM)
    if (const char *e = getenv("B2K_DEC_RS_CAPS")) {          // tuning knob: "tokens,arcs,worklist" (powers of two <= 8192; 0,0,0 = off)
      int a = 0, b = 0, c = 0;
      if (sscanf(e, "%d,%d,%d", &a, &b, &c) == 3 && a >= 0 && b >= 0 && c >= 0 && a <= 8192 && b <= 8192 && c <= 8192) {
        // powers of two only: the kernel's key sort pads the replay worklist to a power of two inside
        // the (tokens + arcs) * 8-byte area
        auto pow2_floor = [](int v) { int q = 1; while (q * 2 <= v) q *= 2; return v > 0 ? q : 0; };
        p.rs_rcap = pow2_floor(a); p.rs_ecap = pow2_floor(b); p.rs_qcap = pow2_floor(c);
        if (!a || !b || !c) { p.rs_rcap = p.rs_ecap = p.rs_qcap = 0; }
      }
    }
    A(p.x_xb, 4 * nl * p.max_tpf, 0);
    A(p.x_rec, sizeof(int4) * nl * p.max_tpf, 0);
    A(p.x_newseq, 4 * nl * p.max_tpf, 0);
    A(p.x_adj, sizeof(int4) * nl * p.adj_cap, 0);
    A(p.x_adjo, 4 * nl * p.adj_cap, 0);
    {
      const char *e = getenv("B2K_DEC_V1");
      // the second generation packs (key, key, creation index) into 64 bits for the replay's initial worklist
      d->use_v2 = !(e && atoi(e) != 0) && p.pos_cap <= (1 << 20) && p.max_tpf <= (1 << 17);
      p.v2_l1_shift = 3;      // measured (profiles/r02_decoder_history.md, r2j-r2k): 552 ms at 1x / 2x, 472 ms at 4x, 463 ms at 8x the HashList size
      if (const char *lx = getenv("B2K_DEC_L1X")) p.v2_l1_shift = std::max(0, std::min(6, atoi(lx)));   // level-1 window = 2^x times the HashList size
    }
    if (d->use_v2) {
      p.v2_hw_len = ((p.pos_cap + 2 * p.max_tpf + 64 + 63) / 64) * 64;
      A(p.v2_trec, sizeof(int4) * nl * p.max_tpf, 0);
      A(p.v2_tok4, sizeof(int4) * nl * p.max_tpf, 0);
      A(p.v2_c0, 4 * nl * p.max_tpf, 0);
      A(p.v2_adjc, 4 * nl * p.adj_cap, 0);
      A(p.v2_qstamp, 4 * nl * p.hash_size, 0);
      A(p.v2_hw, 4 * nl * (size_t)p.v2_hw_len, 0);
      A(p.v2_pf, 4 * nl * 2 * p.max_tpf, 0);
    }
    {
      std::vector<int4> empty((size_t)p.hc_cap, make_int4(0x7fffffff, 0, 0, 0));
      for (size_t l = 0; l < nl; l++)
        B2K_CUDA_CHECK(cudaMemcpy(p.x_bk + l * p.hc_cap, empty.data(), sizeof(int4) * empty.size(), cudaMemcpyHostToDevice));
    }
  }
  p.cap_ls = cfg->max_lattice_states > 0 ? cfg->max_lattice_states : 131072;
  p.cap_la = cfg->max_lattice_arcs > 0 ? cfg->max_lattice_arcs : 262144;
  p.cap_lf = p.max_tpf;
  A(p.lat_states, sizeof(int4) * nc * p.cap_ls, 0);
  A(p.lat_arcs, sizeof(int4) * nc * p.cap_la, 0);
  A(p.lat_arcw, sizeof(float2) * nc * p.cap_la, 0);
  A(p.lat_finals, sizeof(int2) * nc * p.cap_lf, 0);
  A(d->d_pack_ch, 4 * nc, 0);
  A(d->d_pack_offs, 8 * 3 * (nc + 1), 0);
  A(d->d_lane_channel, 4 * nlanes_sz, 0);
  A(d->d_lane_ll, sizeof(float *) * nlanes_sz, 0);
  A(d->d_lane_nframes, 4 * nlanes_sz, 0);
  A(d->d_lane_counter, 4, 0);
  p.lane_counter = d->d_lane_counter;
  p.num_pdfs = fst->num_pdfs_seen;
#undef A
  // hash init: key = EMPTY, cost = +inf (ord), tok = 0, stamp = 0
  {
    std::vector<int4> init((size_t)p.hash_size, make_int4(B2K_HASH_EMPTY, (int)B2K_INF_ORD, cfg->reference_order ? 0x7fffffff : 0, 0));
    for (int l = 0; l < d->nslots; l++)
      B2K_CUDA_CHECK(cudaMemcpy(p.hash + (size_t)l * p.hash_size, init.data(),
                                sizeof(int4) * init.size(), cudaMemcpyHostToDevice));
  }
  // channels start un-initialised
  {
    std::vector<ChanState> cs(nc);
    memset(cs.data(), 0, sizeof(ChanState) * nc);
    for (auto &c : cs) c.frames_decoded = -1;
    B2K_CUDA_CHECK(cudaMemcpy(p.chan, cs.data(), sizeof(ChanState) * nc, cudaMemcpyHostToDevice));
  }
  B2K_CUDA_CHECK(cudaMallocHost((void **)&d->h_chan, sizeof(ChanState) * nc));
  B2K_CUDA_CHECK(cudaMallocHost((void **)&d->h_lane_channel, 4 * nlanes_sz));
  B2K_CUDA_CHECK(cudaMallocHost((void **)&d->h_lane_ll, sizeof(float *) * nlanes_sz));
  B2K_CUDA_CHECK(cudaMallocHost((void **)&d->h_lane_nframes, 4 * nlanes_sz));
  B2K_CUDA_CHECK(cudaEventCreateWithFlags(&d->staging_free, cudaEventDisableTiming));
  p.lane_channel = d->d_lane_channel;
  p.lane_loglikes = d->d_lane_ll;
  p.lane_nframes = d->d_lane_nframes;
  return B2K_OK;
}

int b2k_dec_destroy(b2k_dec *d) {
  if (!d) return B2K_OK;
  cudaDeviceSynchronize();
  for (void *p : d->allocs) cudaFree(p);
  if (d->h_lane_channel) cudaFreeHost(d->h_lane_channel);
  if (d->h_lane_ll) cudaFreeHost(d->h_lane_ll);
  if (d->h_lane_nframes) cudaFreeHost(d->h_lane_nframes);
  if (d->h_chan) cudaFreeHost(d->h_chan);
  if (d->h_pack) cudaFreeHost(d->h_pack);
  if (d->d_pack) cudaFree(d->d_pack);
  if (d->h_bp) cudaFreeHost(d->h_bp);
  if (d->d_bp) cudaFree(d->d_bp);
  if (d->staging_free) cudaEventDestroy(d->staging_free);
  delete d;
  return B2K_OK;
}

#define DEC_THREADS 256

// CTA width of the reference-order kernel.  The kernel is bound by the latency/throughput of
// scattered L2/HBM accesses, so a lane wants as many threads as the SM can give it: one
// 1024-thread CTA per SM while the batch fits one wave, two 512-thread CTAs per SM beyond
// that (measured on B200 at 592 lanes: 4x256 790 ms, 1x1024 758 ms, 2x512 743 ms).
static int dec_threads(const b2k_dec *d, int nlanes) {
  if (d->threads_override > 0) return d->threads_override;
  return nlanes <= d->num_sms ? 1024 : 512;
}

static size_t exact_smem_bytes(const DecParams &p) {
  if (p.rs_rcap == 0) return 0;
  return sizeof(float2) * ((size_t)p.rs_rcap + p.rs_ecap) + sizeof(unsigned short) * ((size_t)p.rs_rcap + p.rs_qcap) + 16;
}

// every per-lane launch is persistent: zero the lane counter, at most `max_grid` CTAs
static int begin_persistent(const b2k_dec *d, DecParams &p, int n, cudaStream_t st) {
  p.n_lanes = n;
  B2K_CUDA_CHECK(cudaMemsetAsync(d->d_lane_counter, 0, 4, st));
  return B2K_OK;
}

extern "C++" {
template <int T, bool PROF>
static int launch_exact_t(const b2k_dec *d, DecParams p, int n, int ctas_per_sm, cudaStream_t st) {
  static size_t configured = 0;
  static size_t static_smem = 0;
  if (!static_smem) {
    cudaFuncAttributes fa;
    B2K_CUDA_CHECK(cudaFuncGetAttributes(&fa, dec_advance_exact_kernel<T, PROF>));
    static_smem = fa.sharedSizeBytes;
  }
  p.rs_bytes = (int32_t)exact_smem_bytes(p);
  size_t smem = (size_t)p.rs_bytes;
  // the log-likelihood row goes to shared memory when it fits beside the replay arrays at this occupancy
  const size_t ll_bytes = ((size_t)p.num_pdfs * 4 + 15) / 16 * 16;
  const size_t per_cta = (size_t)227 * 1024 / (size_t)ctas_per_sm - 1024;
  p.ll_smem = (!d->ll_smem_off && p.num_pdfs > 0 && !p.do_init && p.row_stride >= p.num_pdfs &&
               static_smem + smem + ll_bytes <= per_cta) ? 1 : 0;
  if (p.ll_smem) smem += ll_bytes;
  if (smem > configured) {                                   // static + dynamic may exceed 48 KB: always opt in
    B2K_CUDA_CHECK(cudaFuncSetAttribute(dec_advance_exact_kernel<T, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int grid = std::min(n, std::min(d->nslots, d->num_sms * ctas_per_sm));
  dec_advance_exact_kernel<T, PROF><<<grid, T, smem, st>>>(p);
  return B2K_OK;
}
}  // extern "C++"

extern "C++" {
template <int T, bool PROF, int IT = 4>
static int launch_v2_t(const b2k_dec *d, DecParams p, int n, int ctas_per_sm, cudaStream_t st) {
  static size_t configured = 0;
  static size_t static_smem = 0;
  if (!static_smem) {
    cudaFuncAttributes fa;
    B2K_CUDA_CHECK(cudaFuncGetAttributes(&fa, dec_advance_v2_kernel<T, PROF, IT>));
    static_smem = fa.sharedSizeBytes;
  }
  p.rs_bytes = (int32_t)exact_smem_bytes(p);
  // the log-likelihood row shares the replay arrays' space (idle during the expansion): it only costs shared memory
  // when it is larger than them
  const size_t ll_bytes = ((size_t)p.num_pdfs * 4 + 15) / 16 * 16;
  const size_t per_cta = (size_t)227 * 1024 / (size_t)ctas_per_sm - 1024;
  size_t smem = (size_t)p.rs_bytes;
  p.ll_smem = (!d->ll_smem_off && p.num_pdfs > 0 && !p.do_init && p.row_stride >= p.num_pdfs &&
               static_smem + std::max(smem, ll_bytes) <= per_cta) ? 1 : 0;
  if (p.ll_smem) smem = std::max(smem, ll_bytes);
  p.cid_smem = d->cid_smem_off ? 0 : 1;
  p.par_walk = d->par_walk_off ? 0 : 1;
  p.hash_prefetch = (getenv("B2K_DEC_PREFETCH") && atoi(getenv("B2K_DEC_PREFETCH")) == 0) ? 0 : 1;   // measured 427 -> 422 ms (r2q)
  if (smem > configured) {
    B2K_CUDA_CHECK(cudaFuncSetAttribute(dec_advance_v2_kernel<T, PROF, IT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  int grid = std::min(n, std::min(d->nslots, d->num_sms * ctas_per_sm));
  if (d->grid_cap > 0) grid = std::min(grid, d->grid_cap);
  dec_advance_v2_kernel<T, PROF, IT><<<grid, T, smem, st>>>(p);
  return B2K_OK;
}
}  // extern "C++"

static int launch_exact(const b2k_dec *d, DecParams p, int n, cudaStream_t st) {
  int rc = begin_persistent(d, p, n, st);
  if (rc) return rc;
  const int threads = dec_threads(d, n);
  if (d->use_v2) {
    if (d->prof) {
      if (threads == 1024) return launch_v2_t<1024, true>(d, p, n, 1, st);
      if (threads == 512) return d->arcs_per_thread == 3 ? launch_v2_t<512, true, 3>(d, p, n, 2, st) : launch_v2_t<512, true>(d, p, n, 2, st);
      return launch_v2_t<256, true>(d, p, n, d->ctas_override > 0 ? d->ctas_override : 2, st);
    }
    if (threads == 1024) return launch_v2_t<1024, false>(d, p, n, 1, st);
    if (threads == 512) {
      if (d->arcs_per_thread == 2) return launch_v2_t<512, false, 2>(d, p, n, 2, st);   // B2K_DEC_IT: arcs per thread and admission round
      if (d->arcs_per_thread == 3) return launch_v2_t<512, false, 3>(d, p, n, 2, st);
      return launch_v2_t<512, false>(d, p, n, 2, st);
    }
    if (threads == 128) return launch_v2_t<128, false>(d, p, n, 2, st);
    return launch_v2_t<256, false>(d, p, n, d->ctas_override > 0 ? d->ctas_override : 2, st);
  }
  if (d->prof) {
    if (threads == 1024) return launch_exact_t<1024, true>(d, p, n, 1, st);
    if (threads == 512) return launch_exact_t<512, true>(d, p, n, 2, st);
    return launch_exact_t<256, true>(d, p, n, 2, st);
  }
  if (threads == 1024) return launch_exact_t<1024, false>(d, p, n, 1, st);
  if (threads == 512) return launch_exact_t<512, false>(d, p, n, 2, st);
  if (threads == 128) return launch_exact_t<128, false>(d, p, n, 2, st);
  return launch_exact_t<256, false>(d, p, n, 2, st);
}

static int stage_lanes(b2k_dec *d, const int32_t *channels, const float *const *lls,
                       const int32_t *nframes, int n, cudaStream_t st) {
  if (n <= 0 || n > d->nlanes) return set_error(B2K_ERR_INVALID, "number of lanes out of range");
  // pinned staging is reused: wait until the previous launch consumed it
  B2K_CUDA_CHECK(cudaEventSynchronize(d->staging_free));
  for (int i = 0; i < n; i++) {
    if (channels[i] < 0 || channels[i] >= d->nchannels) return set_error(B2K_ERR_INVALID, "bad channel id");
    d->h_lane_channel[i] = channels[i];
    d->h_lane_ll[i] = lls ? lls[i] : nullptr;
    d->h_lane_nframes[i] = nframes ? nframes[i] : 1;
  }
  B2K_CUDA_CHECK(cudaMemcpyAsync(d->d_lane_channel, d->h_lane_channel, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
  B2K_CUDA_CHECK(cudaMemcpyAsync((void *)d->d_lane_ll, d->h_lane_ll, sizeof(float *) * (size_t)n, cudaMemcpyHostToDevice, st));
  B2K_CUDA_CHECK(cudaMemcpyAsync(d->d_lane_nframes, d->h_lane_nframes, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
  B2K_CUDA_CHECK(cudaEventRecord(d->staging_free, st));
  return B2K_OK;
}

int b2k_dec_init_decoding(b2k_dec *d, const int32_t *channels, int32_t n, void *stream) {
  if (!d || !channels) return set_error(B2K_ERR_INVALID, "b2k_dec_init_decoding: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = stage_lanes(d, channels, nullptr, nullptr, n, st);          // (validates the channel ids)
  if (rc) return rc;
  DecParams p = d->p;
  dec_reset_channels_kernel<<<(n + 127) / 128, 128, 0, st>>>(p, n);       // channel state as b2k_dec_create leaves it
  B2K_LAUNCH_CHECK();
  p.do_init = 1;
  if (d->cfg.reference_order) { if ((rc = launch_exact(d, p, n, st))) return rc; }
  else {
    if ((rc = begin_persistent(d, p, n, st))) return rc;
    dec_advance_kernel<DEC_THREADS><<<std::min(n, d->nslots), DEC_THREADS, 0, st>>>(p);
  }
  B2K_LAUNCH_CHECK();
  return B2K_OK;
}

int b2k_dec_advance_decoding_frames(b2k_dec *d, const int32_t *channels,
                                    const float *const *d_loglikes, const int32_t *num_frames,
                                    int32_t row_stride, int32_t n, void *stream) {
  if (!d || !channels || !d_loglikes) return set_error(B2K_ERR_INVALID, "b2k_dec_advance_decoding: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = stage_lanes(d, channels, d_loglikes, num_frames, n, st);
  if (rc) return rc;
  DecParams p = d->p;
  p.do_init = 0;
  p.row_stride = row_stride;
  if (d->cfg.reference_order) { if ((rc = launch_exact(d, p, n, st))) return rc; }
  else {
    if ((rc = begin_persistent(d, p, n, st))) return rc;
    dec_advance_kernel<DEC_THREADS><<<std::min(n, d->nslots), DEC_THREADS, 0, st>>>(p);
  }
  B2K_LAUNCH_CHECK();
  return B2K_OK;
}

int b2k_dec_advance_decoding(b2k_dec *d, const int32_t *channels, const float *const *d_loglikes,
                             int32_t n, void *stream) {
  return b2k_dec_advance_decoding_frames(d, channels, d_loglikes, nullptr, 0, n, stream);
}

int b2k_dec_finalize_decoding(b2k_dec *d, const int32_t *channels, int32_t n, void *stream) {
  if (!d || !channels) return set_error(B2K_ERR_INVALID, "b2k_dec_finalize_decoding: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = stage_lanes(d, channels, nullptr, nullptr, n, st);
  if (rc) return rc;
  DecParams p = d->p;
  {
    const int fin_threads = d->fin_threads;
    if ((rc = begin_persistent(d, p, n, st))) return rc;
    const int grid = std::min(n, d->nslots);                 // one resident CTA per scratch slot
    p.fin_scap = 0;
    if (fin_threads == 1024) {
      // one CTA per SM: 16 K tokens per list in shared memory (2 x 64 KB + three bitmaps, inside the 164 KB carve-out)
      static bool configured = false;
      const int scap = d->fin_smem_off ? 0 : 16384;
      const size_t smem = (size_t)scap * 8 + 3 * (size_t)((scap + 31) / 32) * 4;
      if (!configured) {
        B2K_CUDA_CHECK(cudaFuncSetAttribute(dec_finalize_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)16384 * 8 + 3 * 512 * 4)));
        configured = true;
      }
      p.fin_scap = scap;
      dec_finalize_kernel<1024><<<std::min(grid, d->num_sms), 1024, smem, st>>>(p);
    }
    else if (fin_threads == 512) dec_finalize_kernel<512><<<grid, 512, 0, st>>>(p);
    else dec_finalize_kernel<256><<<grid, 256, 0, st>>>(p);
  }
  B2K_LAUNCH_CHECK();
  return B2K_OK;
}

static int read_chan(b2k_dec *d, int ch, ChanState *out) {
  if (!d) return set_error(B2K_ERR_INVALID, "null decoder handle");
  if (ch < 0 || ch >= d->nchannels) return set_error(B2K_ERR_INVALID, "bad channel id");
  B2K_CUDA_CHECK(cudaDeviceSynchronize());
  B2K_CUDA_CHECK(cudaMemcpy(out, &d->p.chan[ch], sizeof(ChanState), cudaMemcpyDeviceToHost));
  return B2K_OK;
}

int b2k_dec_num_frames_decoded(b2k_dec *d, int32_t channel, int32_t *out) {
  if (!out) return set_error(B2K_ERR_INVALID, "b2k_dec_num_frames_decoded: bad args");
  ChanState cs;
  int rc = read_chan(d, channel, &cs);
  if (rc) return rc;
  *out = cs.frames_decoded;
  return B2K_OK;
}

int b2k_dec_channel_info(b2k_dec *d, int32_t channel, int64_t info[32]) {
  if (!info) return set_error(B2K_ERR_INVALID, "b2k_dec_channel_info: bad args");
  ChanState cs;
  int rc = read_chan(d, channel, &cs);
  if (rc) return rc;
  memset(info, 0, sizeof(int64_t) * 32);
  info[0] = cs.status; info[1] = cs.frames_decoded; info[2] = cs.ntok; info[3] = cs.nlink;
  info[4] = (int64_t)cs.arcs_e; info[5] = (int64_t)cs.arcs_ne; info[6] = cs.lat_states;
  info[7] = cs.lat_arcs; info[8] = cs.lat_finals; info[9] = cs.finalized; info[10] = cs.any_final; info[11] = cs.err_line;
  for (int k = 0; k < 16; k++) info[16 + k] = (int64_t)cs.prof[k];
  return B2K_OK;
}

// Asynchronous packing for pipelined read-back (kaldi_b200/csrc/pipeline.cu: submit / collect): the lattices of n
// finalized channels are packed into d_buf entirely on the device; the host later copies the header, learns the sizes,
// copies the body and unpacks it with b2k_dec_unpack_lattices.
int64_t b2k_dec_pack_header_bytes(int32_t n) { return n < 0 ? 0 : (int64_t)b2k::pack_header_bytes(n); }

int b2k_dec_pack_lattices_async(b2k_dec *d, const int32_t *d_channels, int32_t n, void *d_buf, int64_t cap_bytes, void *stream) {
  if (!d || !d_channels || n <= 0 || n > d->nchannels || !d_buf || cap_bytes < (int64_t)b2k::pack_header_bytes(n))
    return set_error(B2K_ERR_INVALID, "b2k_dec_pack_lattices_async: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  dec_pack_header_kernel<<<1, 1024, 0, st>>>(d->p, d_channels, n, (int64_t *)d_buf, cap_bytes);
  B2K_LAUNCH_CHECK();
  dec_pack_body_kernel<<<dim3(8, n), 256, 0, st>>>(d->p, d_channels, n, (const int64_t *)d_buf, (char *)d_buf);
  B2K_LAUNCH_CHECK();
  return B2K_OK;
}

// h_buf: a host copy of the packed buffer (header, then body).  out arrays must hold the totals the header states
// (query them with out->state_frame == NULL).
int b2k_dec_unpack_lattices(const void *h_buf, int32_t n, b2k_raw_lattice *out, int64_t *state_offs, int64_t *arc_offs, int64_t *final_offs) {
  if (!h_buf || n <= 0 || !out) return set_error(B2K_ERR_INVALID, "b2k_dec_unpack_lattices: bad args");
  const int64_t *hdr = (const int64_t *)h_buf;
  const int64_t *tail = hdr + 3 * ((size_t)n + 1);
  if (tail[2] != n) return set_error(B2K_ERR_INVALID, "b2k_dec_unpack_lattices: the buffer was packed for another channel count");
  if (tail[0] != B2K_OK) {
    char msg[240];
    if (tail[3] >= 0)
      snprintf(msg, sizeof(msg), "packed lattices carry status %lld: entry %lld of the channel list is in error (raised at decoder.cu:%lld; 0 = not finalized)",
               (long long)tail[0], (long long)(tail[3] & 0xffffffffll), (long long)(tail[3] >> 32));
    else
      snprintf(msg, sizeof(msg), "packed lattices carry status %lld: the buffer was too small (%lld bytes needed)", (long long)tail[0], (long long)tail[1]);
    return set_error((int)tail[0], msg);
  }
  const int64_t ns = hdr[n], na = hdr[(n + 1) + n], nf = hdr[2 * (n + 1) + n];
  if (state_offs) for (int i = 0; i <= n; i++) state_offs[i] = hdr[i];
  if (arc_offs) for (int i = 0; i <= n; i++) arc_offs[i] = hdr[(n + 1) + i];
  if (final_offs) for (int i = 0; i <= n; i++) final_offs[i] = hdr[2 * (n + 1) + i];
  if (!out->state_frame) { out->num_states = ns; out->num_arcs = na; out->num_finals = nf; return B2K_OK; }
  if (out->num_states < ns || out->num_arcs < na || out->num_finals < nf)
    return set_error(B2K_ERR_INVALID, "b2k_dec_unpack_lattices: output buffers too small");
  const int4 *hs = (const int4 *)((const char *)h_buf + b2k::pack_header_bytes(n));
  const int4 *ha = hs + ns;
  const float2 *hw = (const float2 *)(ha + na);
  const int2 *hf = (const int2 *)(hw + na);
  for (int64_t i = 0; i < ns; i++) {
    out->state_frame[i] = hs[i].x; out->state_hclg[i] = hs[i].y;
    memcpy(&out->state_tot_cost[i], &hs[i].z, 4); memcpy(&out->state_extra_cost[i], &hs[i].w, 4);
  }
  for (int64_t i = 0; i < na; i++) {
    out->arc_src[i] = ha[i].x; out->arc_dst[i] = ha[i].y; out->arc_ilabel[i] = ha[i].z; out->arc_olabel[i] = ha[i].w;
    out->arc_graph_cost[i] = hw[i].x; out->arc_acoustic_cost[i] = hw[i].y;
  }
  for (int64_t i = 0; i < nf; i++) { out->final_state[i] = hf[i].x; memcpy(&out->final_cost[i], &hf[i].y, 4); }
  out->num_states = ns; out->num_arcs = na; out->num_finals = nf;
  return B2K_OK;
}

// Batched read-back: sizes first (one D2H of the channel states), then one
// pack kernel and one D2H for all channels.
int b2k_dec_get_raw_lattices(b2k_dec *d, const int32_t *channels, int32_t n, b2k_raw_lattice *out,
                             int64_t *state_offs, int64_t *arc_offs, int64_t *final_offs, void *stream) {
  if (!d || !channels || n <= 0 || n > d->nchannels || !out || !state_offs || !arc_offs || !final_offs)
    return set_error(B2K_ERR_INVALID, "b2k_dec_get_raw_lattices: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  B2K_CUDA_CHECK(cudaMemcpyAsync(d->h_chan, d->p.chan, sizeof(ChanState) * d->nchannels, cudaMemcpyDeviceToHost, st));
  B2K_CUDA_CHECK(cudaStreamSynchronize(st));
  std::vector<int64_t> offs(3 * (size_t)(n + 1), 0);
  for (int i = 0; i < n; i++) {
    int ch = channels[i];
    if (ch < 0 || ch >= d->nchannels) return set_error(B2K_ERR_INVALID, "bad channel id");
    const ChanState &cs = d->h_chan[ch];
    if (cs.status != B2K_OK) {
      char msg[160];
      snprintf(msg, sizeof(msg), "channel %d is in error state %d, raised at decoder.cu:%d (capacity overflow? see b2k_dec_cfg)", ch, cs.status, cs.err_line);
      return set_error(cs.status, msg);
    }
    if (!cs.finalized) return set_error(B2K_ERR_STATE, "call b2k_dec_finalize_decoding first");
    offs[i + 1] = offs[i] + cs.lat_states;
    offs[(n + 1) + i + 1] = offs[(n + 1) + i] + cs.lat_arcs;
    offs[2 * (n + 1) + i + 1] = offs[2 * (n + 1) + i] + cs.lat_finals;
  }
  const int64_t ns = offs[n], na = offs[(n + 1) + n], nf = offs[2 * (n + 1) + n];
  for (int i = 0; i <= n; i++) { state_offs[i] = offs[i]; arc_offs[i] = offs[(n + 1) + i]; final_offs[i] = offs[2 * (n + 1) + i]; }
  if (!out->state_frame) { out->num_states = ns; out->num_arcs = na; out->num_finals = nf; return B2K_OK; }
  if (out->num_states < ns || out->num_arcs < na || out->num_finals < nf)
    return set_error(B2K_ERR_INVALID, "b2k_dec_get_raw_lattices: output buffers too small");
  size_t bytes = (size_t)ns * 16 + (size_t)na * 24 + (size_t)nf * 8 + 64;
  if (bytes > d->d_pack_bytes) {
    if (d->d_pack) cudaFree(d->d_pack);
    if (d->h_pack) cudaFreeHost(d->h_pack);
    size_t cap = bytes + bytes / 4;
    B2K_CUDA_CHECK(cudaMalloc((void **)&d->d_pack, cap));
    B2K_CUDA_CHECK(cudaMallocHost((void **)&d->h_pack, cap));
    d->d_pack_bytes = d->h_pack_bytes = cap;
  }
  int4 *o_states = (int4 *)d->d_pack;
  int4 *o_arcs = o_states + ns;
  float2 *o_arcw = (float2 *)(o_arcs + na);
  int2 *o_finals = (int2 *)(o_arcw + na);
  B2K_CUDA_CHECK(cudaMemcpyAsync(d->d_pack_ch, channels, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
  B2K_CUDA_CHECK(cudaMemcpyAsync(d->d_pack_offs, offs.data(), 8 * offs.size(), cudaMemcpyHostToDevice, st));
  dec_pack_kernel<<<dim3(8, n), 256, 0, st>>>(d->p, d->d_pack_ch, d->d_pack_offs, n, o_states, o_arcs, o_arcw, o_finals);
  B2K_LAUNCH_CHECK();
  B2K_CUDA_CHECK(cudaMemcpyAsync(d->h_pack, d->d_pack, bytes - 64, cudaMemcpyDeviceToHost, st));
  B2K_CUDA_CHECK(cudaStreamSynchronize(st));
  const int4 *hs = (const int4 *)d->h_pack;
  const int4 *ha = hs + ns;
  const float2 *hw = (const float2 *)(ha + na);
  const int2 *hf = (const int2 *)(hw + na);
  for (int64_t i = 0; i < ns; i++) {
    out->state_frame[i] = hs[i].x; out->state_hclg[i] = hs[i].y;
    memcpy(&out->state_tot_cost[i], &hs[i].z, 4); memcpy(&out->state_extra_cost[i], &hs[i].w, 4);
  }
  for (int64_t i = 0; i < na; i++) {
    out->arc_src[i] = ha[i].x; out->arc_dst[i] = ha[i].y; out->arc_ilabel[i] = ha[i].z; out->arc_olabel[i] = ha[i].w;
    out->arc_graph_cost[i] = hw[i].x; out->arc_acoustic_cost[i] = hw[i].y;
  }
  for (int64_t i = 0; i < nf; i++) { out->final_state[i] = hf[i].x; memcpy(&out->final_cost[i], &hf[i].y, 4); }
  out->num_states = ns; out->num_arcs = na; out->num_finals = nf;
  return B2K_OK;
}

int b2k_dec_get_raw_lattice(b2k_dec *d, int32_t channel, b2k_raw_lattice *out, void *stream) {
  int64_t so[2], ao[2], fo[2];
  return b2k_dec_get_raw_lattices(d, &channel, 1, out, so, ao, fo, stream);
}

int b2k_dec_debug_frame(b2k_dec *d, int32_t channel, int32_t frame_plus_one, int32_t *tok_state,
                        float *tok_cost, int64_t *ntok, int32_t *links7, int64_t *nlink,
                        int64_t cap_tok, int64_t cap_link) {
  ChanState cs;
  int rc = read_chan(d, channel, &cs);
  if (rc) return rc;
  if (frame_plus_one < 0 || frame_plus_one > cs.frames_decoded) return set_error(B2K_ERR_INVALID, "frame out of range");
  const DecParams &p = d->p;
  size_t fo = (size_t)channel * (p.max_frames + 2);
  int32_t tb[2], lb[2], le;
  B2K_CUDA_CHECK(cudaMemcpy(tb, p.frame_tok_begin + fo + frame_plus_one, 8, cudaMemcpyDeviceToHost));
  B2K_CUDA_CHECK(cudaMemcpy(lb, p.frame_link_begin + fo + frame_plus_one, 8, cudaMemcpyDeviceToHost));
  B2K_CUDA_CHECK(cudaMemcpy(&le, p.frame_link_eps + fo + frame_plus_one, 4, cudaMemcpyDeviceToHost));
  int64_t nt = tb[1] - tb[0], nl = lb[1] - lb[0];
  *ntok = nt; *nlink = nl;
  if (!tok_state) return B2K_OK;
  if (nt > cap_tok || nl > cap_link) return set_error(B2K_ERR_INVALID, "debug buffers too small");
  const int32_t *ts = p.tok_state + (size_t)channel * p.max_tokens;
  const float *tc = p.tok_cost + (size_t)channel * p.max_tokens;
  B2K_CUDA_CHECK(cudaMemcpy(tok_state, ts + tb[0], nt * 4, cudaMemcpyDeviceToHost));
  B2K_CUDA_CHECK(cudaMemcpy(tok_cost, tc + tb[0], nt * 4, cudaMemcpyDeviceToHost));
  std::vector<int4> lk(nl);
  B2K_CUDA_CHECK(cudaMemcpy(lk.data(), p.links + (size_t)channel * p.max_links + lb[0], nl * sizeof(int4), cudaMemcpyDeviceToHost));
  // previous list's states for emitting-link sources
  int32_t ptb = 0;
  std::vector<int32_t> prev_states;
  if (frame_plus_one > 0) {
    B2K_CUDA_CHECK(cudaMemcpy(&ptb, p.frame_tok_begin + fo + frame_plus_one - 1, 4, cudaMemcpyDeviceToHost));
    prev_states.resize(tb[0] - ptb);
    B2K_CUDA_CHECK(cudaMemcpy(prev_states.data(), ts + ptb, (size_t)(tb[0] - ptb) * 4, cudaMemcpyDeviceToHost));
  }
  bool bad = false;
  auto state_of = [&](int32_t tok) -> int32_t {
    if (tok >= tb[0] && tok < tb[1]) return tok_state[tok - tb[0]];
    if (tok >= ptb && tok < tb[0]) return prev_states[tok - ptb];
    bad = true;
    return -1;
  };
  const int64_t n_e = (int64_t)d->fst->h_e.size(), n_ne = (int64_t)d->fst->h_ne.size();
  for (int64_t i = 0; i < nl; i++) {
    int4 l = lk[i];
    int32_t *r = links7 + i * 7;
    r[0] = state_of(l.x); r[1] = state_of(l.y);
    const int32_t zi = (int32_t)((uint32_t)l.z & B2K_ARC_MASK);
    if ((uint32_t)l.z & B2K_EPS_FLAG) {
      int64_t a = zi;
      if (a >= n_ne) { bad = true; continue; }
      int4 arc = d->fst->h_ne[a];
      r[2] = 0; r[3] = arc.z & 0x7fffffff; r[4] = arc.y; r[5] = 0; r[6] = 1;
    } else {
      if (zi >= n_e) { bad = true; continue; }
      int4 arc = d->fst->h_e[zi];
      r[2] = d->fst->h_eil[zi]; r[3] = arc.w; r[4] = arc.y; r[5] = l.w; r[6] = 0;
    }
    if (bad) {
      char buf[256];
      snprintf(buf, sizeof(buf), "link %lld of frame %d = {%d,%d,%d,%d}; tok ranges prev [%d,%d) cur [%d,%d) links [%d,%d)",
               (long long)i, frame_plus_one, l.x, l.y, l.z, l.w, ptb, tb[0], tb[0], tb[1], lb[0], lb[1]);
      return set_error(B2K_ERR_STATE, "inconsistent link record", buf);
    }
  }
  (void)le;
  return B2K_OK;
}

int b2k_dec_frame_info(b2k_dec *d, int32_t channel, float *cutoff, float *cost_offset,
                       int32_t *ntoks, int32_t cap) {
  ChanState cs;
  int rc = read_chan(d, channel, &cs);
  if (rc) return rc;
  int T = std::min(cap, cs.frames_decoded);
  if (T <= 0) return B2K_OK;
  const DecParams &p = d->p;
  if (cutoff) B2K_CUDA_CHECK(cudaMemcpy(cutoff, p.frame_cutoff + (size_t)channel * (p.max_frames + 1), 4 * (size_t)T, cudaMemcpyDeviceToHost));
  if (cost_offset) B2K_CUDA_CHECK(cudaMemcpy(cost_offset, p.frame_cost_offset + (size_t)channel * (p.max_frames + 1), 4 * (size_t)T, cudaMemcpyDeviceToHost));
  if (ntoks) {
    std::vector<int32_t> tb(T + 2);
    B2K_CUDA_CHECK(cudaMemcpy(tb.data(), p.frame_tok_begin + (size_t)channel * (p.max_frames + 2), 4 * (size_t)(T + 2), cudaMemcpyDeviceToHost));
    for (int f = 0; f < T; f++) ntoks[f] = tb[f + 2] - tb[f + 1];
  }
  return B2K_OK;
}

int b2k_dec_best_path(b2k_dec *d, const int32_t *channels, int32_t n, int32_t use_final_probs, int32_t cap,
                      int32_t *ilabels, int32_t *olabels, float *graph_costs, float *acoustic_costs, int32_t *arc_frame,
                      int32_t *arc_state, b2k_best_path_info *info, void *stream) {
  if (!d || !channels || n <= 0 || cap <= 0 || !info) return set_error(B2K_ERR_INVALID, "b2k_dec_best_path: bad args");
  for (int i = 0; i < n; i++)
    if (channels[i] < 0 || channels[i] >= d->nchannels) return set_error(B2K_ERR_INVALID, "b2k_dec_best_path: bad channel id");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t hdr_b = sizeof(BestPathHdr) * (size_t)n, arc_b = sizeof(int4) * (size_t)n * cap, wh_b = sizeof(int2) * (size_t)n * cap;
  const size_t ch_b = ((sizeof(int32_t) * (size_t)n + 15) / 16) * 16;
  const size_t need = ch_b + hdr_b + arc_b + wh_b;
  if (need > d->d_bp_bytes) {
    B2K_CUDA_CHECK(cudaStreamSynchronize(st));
    if (d->d_bp) cudaFree(d->d_bp);
    if (d->h_bp) cudaFreeHost(d->h_bp);
    d->d_bp = nullptr; d->h_bp = nullptr; d->d_bp_bytes = d->h_bp_bytes = 0;
    B2K_CUDA_CHECK(cudaMalloc(&d->d_bp, need));
    B2K_CUDA_CHECK(cudaMallocHost(&d->h_bp, need));
    d->d_bp_bytes = d->h_bp_bytes = need;
  }
  memcpy(d->h_bp, channels, sizeof(int32_t) * (size_t)n);
  B2K_CUDA_CHECK(cudaMemcpyAsync(d->d_bp, d->h_bp, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
  BestPathHdr *d_hdr = reinterpret_cast<BestPathHdr *>(d->d_bp + ch_b);
  int4 *d_arcs = reinterpret_cast<int4 *>(d->d_bp + ch_b + hdr_b);
  int2 *d_where = reinterpret_cast<int2 *>(d->d_bp + ch_b + hdr_b + arc_b);
  dec_best_path_kernel<512><<<n, 512, 0, st>>>(d->p, reinterpret_cast<const int32_t *>(d->d_bp), use_final_probs ? 1 : 0, cap, d_arcs, d_where, d_hdr);
  B2K_LAUNCH_CHECK();
  B2K_CUDA_CHECK(cudaMemcpyAsync(d->h_bp + ch_b, d->d_bp + ch_b, hdr_b, cudaMemcpyDeviceToHost, st));
  B2K_CUDA_CHECK(cudaStreamSynchronize(st));
  const BestPathHdr *hh = reinterpret_cast<const BestPathHdr *>(d->h_bp + ch_b);
  int first_bad = 0;
  for (int i = 0; i < n; i++) {
    info[i].status = hh[i].status; info[i].n_arcs = hh[i].n_arcs; info[i].end_state = hh[i].end_state; info[i].num_frames = hh[i].frames;
    info[i].final_cost = hh[i].final_cost; info[i].best_cost = hh[i].best_cost; info[i].final_relative_cost = hh[i].final_relative_cost;
    if (hh[i].status != 0 && !first_bad) first_bad = hh[i].status;
  }
  // the arcs come back end first; hand them out start first
  for (int i = 0; i < n; i++) {
    const int na = std::min(hh[i].n_arcs, cap);
    if (na <= 0) continue;
    const size_t ao = ch_b + hdr_b + sizeof(int4) * (size_t)i * cap, wo = ch_b + hdr_b + arc_b + sizeof(int2) * (size_t)i * cap;
    B2K_CUDA_CHECK(cudaMemcpyAsync(d->h_bp + ao, d->d_bp + ao, sizeof(int4) * (size_t)na, cudaMemcpyDeviceToHost, st));
    B2K_CUDA_CHECK(cudaMemcpyAsync(d->h_bp + wo, d->d_bp + wo, sizeof(int2) * (size_t)na, cudaMemcpyDeviceToHost, st));
  }
  B2K_CUDA_CHECK(cudaStreamSynchronize(st));
  for (int i = 0; i < n; i++) {
    const int na = std::min(hh[i].n_arcs, cap);
    const int4 *a = reinterpret_cast<const int4 *>(d->h_bp + ch_b + hdr_b) + (size_t)i * cap;
    const int2 *w = reinterpret_cast<const int2 *>(d->h_bp + ch_b + hdr_b + arc_b) + (size_t)i * cap;
    for (int k = 0; k < na; k++) {
      const int4 x = a[na - 1 - k];
      const size_t o = (size_t)i * cap + k;
      if (ilabels) ilabels[o] = x.x;
      if (olabels) olabels[o] = x.y;
      if (graph_costs) memcpy(&graph_costs[o], &x.z, 4);
      if (acoustic_costs) memcpy(&acoustic_costs[o], &x.w, 4);
      if (arc_frame) arc_frame[o] = w[na - 1 - k].x;
      if (arc_state) arc_state[o] = w[na - 1 - k].y;
    }
  }
  if (first_bad == B2K_ERR_OVERFLOW) return set_error(B2K_ERR_OVERFLOW, "b2k_dec_best_path: a path is longer than cap arcs");
  if (first_bad) return set_error(first_bad, "b2k_dec_best_path: a channel is not decoding, is in error, or its traceback did not reach the start state");
  return B2K_OK;
}

}  // extern "C"
