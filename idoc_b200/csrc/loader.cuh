// Slab movers shared by the time-sweep kernels.
//
// Every warp streams the per-timestep slabs of its own G problems HBM -> shared memory -> HBM:
//   loads : per-lane cp.async (LDGSTS) of GR-byte chunks, consecutive lanes copying consecutive chunks of one
//           problem's slab (coalesced), optionally double-buffered;
//   stores: LDS + streaming STG of GR-byte chunks with the same lane mapping.
// Two alternatives were built and measured on B200 at n=8, m=4 (profiles/r1_notes.md) and rejected:
//   * 1-D bulk async copies (cp.async.bulk, the TMA engine) on per-warp mbarriers: the per-(problem, array,
//     timestep) segments are 32-512 B and the engine is request-rate bound at that size; the fp64 step went from
//     12.0 ms to 13.6-17.5 ms (commit "wip: multi-TU build, TMA bulk loader experiment");
//   * streaming LDG.128 into registers during the step + STS.128 at its end: 14.5 ms (register pressure).
#pragma once
#include "common.cuh"

namespace idoc {

// G groups per warp, GR-byte chunks, RL rounds of "main" segments, RQ rounds of time-shifted segments
// (segments whose step-t copy reads timestep t+shift and is skipped when that is negative).
// NBULK > 0: that many large segments per group are moved by bulk async copies (one lane per (group, segment)
// pair, completion on a per-warp mbarrier) instead of cp.async; G * NBULK <= 32.
template <int G, int GR, int RL, int RQ, int NBULK = 0>
struct SlabLoader {
  static constexpr int RQ1 = RQ > 0 ? RQ : 1;
  static_assert(G * NBULK <= 32, "one lane per bulk (group, segment) pair");
  CopyPlan<RL, GR> main;
  CopyPlan<RQ1, GR> sh;
  unsigned pidx[G];  // (clamped problem index) * horizon, per group
  int group_bytes, lane, shift;
  // bulk part (per lane: its own pair)
  const char* bgp;
  unsigned bidx;
  int bbytes, bsoff, nbulk_added;
  uint64_t* bars;
  uint32_t phase;

  IDOC_DEVICE void init(long long prob0, long long last, int Th, int group_bytes_, int lane_, uint64_t* bars_ = nullptr) {
    main.reset();
    sh.reset();
    group_bytes = group_bytes_;
    lane = lane_;
    shift = 0;
    bgp = nullptr;
    bbytes = bsoff = nbulk_added = 0;
    bidx = 0;
    bars = bars_;
    phase = 0;
#pragma unroll
    for (int g = 0; g < G; g++) {
      long long p = prob0 + g;
      p = p <= last ? p : last;  // idle groups shadow the last problem (their stores are masked)
      pidx[g] = (unsigned)(p * Th);
      if (NBULK > 0 && lane / NBULK == g) bidx = pidx[g];
    }
    if (NBULK > 0) {
      if (lane == 0) {
        mbar_init(&bars[0], G * NBULK);
        mbar_init(&bars[1], G * NBULK);
        fence_mbar_init();
      }
      __syncwarp();
    }
  }
  // a large segment moved by the TMA engine (call exactly NBULK times, all lanes)
  IDOC_DEVICE void add_bulk(const void* base, int seg_bytes, int smem_off) {
    if (NBULK > 0 && lane < G * NBULK && lane % (NBULK > 0 ? NBULK : 1) == nbulk_added) {
      bgp = static_cast<const char*>(base);
      bbytes = seg_bytes;
      bsoff = (lane / (NBULK > 0 ? NBULK : 1)) * group_bytes + smem_off;
    }
    nbulk_added++;
  }
  IDOC_DEVICE void add(const void* base, int seg_bytes, int sub_off, int sub_bytes, int smem_off, int tshift = 0) {
    if (tshift == 0) {
      main.add(base, seg_bytes, sub_off, sub_bytes, smem_off, lane);
    } else {
      sh.add(base, seg_bytes, sub_off, sub_bytes, smem_off, lane);
      shift = tshift;
    }
  }
  // groups_base32: 32-bit shared address of group 0's slab; buf_off: byte offset of the in-slab buffer
  IDOC_DEVICE void issue(uint32_t groups_base32, int buf_off, int t, int buf = 0) const {
    if (NBULK > 0 && bbytes) {
      mbar_expect_tx(&bars[buf], (uint32_t)bbytes);
      bulk_g2s(groups_base32 + buf_off + bsoff, bgp + (unsigned long long)(bidx + (unsigned)t) * (unsigned)bbytes, (uint32_t)bbytes,
               &bars[buf]);
    }
#pragma unroll
    for (int g = 0; g < G; g++) {
      const unsigned idx = pidx[g] + (unsigned)t;
      const uint32_t slab = groups_base32 + g * group_bytes + buf_off;
      main.load(slab, idx);
      if (RQ > 0 && t + shift >= 0) sh.load(slab, idx + shift);
    }
    cp_async_commit();
  }
  IDOC_DEVICE void wait(int buf = 0) {
    cp_async_wait<0>();
    if (NBULK > 0) {
      mbar_wait(&bars[buf], (phase >> buf) & 1u);
      phase ^= (1u << buf);
    }
  }
};

// shared -> global (all lanes' shared-memory writes must be complete, i.e. a __syncwarp(), before store())
template <int G, int GR, int RS>
struct SlabStorer {
  CopyPlan<RS, GR> st;
  unsigned pidx[G];
  bool act[G];
  int group_bytes, lane;
  IDOC_DEVICE void init(long long prob0, long long last, int Th, int group_bytes_, int lane_) {
    st.reset();
    group_bytes = group_bytes_;
    lane = lane_;
#pragma unroll
    for (int g = 0; g < G; g++) {
      act[g] = prob0 + g <= last;
      pidx[g] = (unsigned)((prob0 + g) * Th);
    }
  }
  IDOC_DEVICE void add(void* base, int seg_bytes, int sub_off, int sub_bytes, int smem_off) {
    st.add(base, seg_bytes, sub_off, sub_bytes, smem_off, lane);
  }
  IDOC_DEVICE void store(const char* groups_base, int buf_off, int t) const {
#pragma unroll
    for (int g = 0; g < G; g++)
      if (act[g]) st.store(groups_base + g * group_bytes + buf_off, pidx[g] + (unsigned)t);
  }
};

}  // namespace idoc
