// Slab movers shared by the time-sweep kernels.
//
// Every warp streams the per-timestep slabs of its own G problems HBM -> shared memory -> HBM:
//   loads : per-lane cp.async (LDGSTS) of GR-byte chunks, consecutive lanes copying consecutive chunks of one
//           problem's slab (coalesced), optionally double-buffered;
//   stores: LDS + streaming STG of GR-byte chunks with the same lane mapping.
//   hybrid: the few large (>= 256 B) segments of a step can instead go by 1-D bulk async copies (the TMA engine).
// Two alternatives were built and measured on B200 at n=8, m=4 (profiles/r1_notes.md) and rejected:
//   * bulk async copies for EVERY segment: the per-(problem, array, timestep) segments are 32-512 B and the
//     engine is request-rate bound at that size; the fp64 step went from 12.0 ms to 13.6-17.5 ms;
//   * streaming LDG.128 into registers during the step + STS.128 at its end: 14.5 ms (register pressure).
#pragma once
#include "common.cuh"

namespace idoc {

// G groups per warp, GR-byte chunks, RL rounds of "main" segments, RQ rounds of time-shifted segments
// (segments whose step-t copy reads timestep t+shift and is skipped when that is negative).
// NBULK > 0: that many large segments per group are moved by bulk async copies (the TMA engine; completion counted in
// bytes on a per-warp, per-buffer mbarrier) instead of cp.async.  The bulk path writes shared memory through the async
// proxy and costs no LSU wavefronts; it pays off for the >= 256-byte segments only (the engine is request-rate bound
// below that).  All G * NBULK copies of a step are issued by ONE elected lane from warp-uniform addresses (the caller
// must pass warp-uniform prob0 / groups_base32): per-lane issue makes the compiler serialise the warp over a
// ~9-instruction loop per copy (UBLKCP takes its operands from uniform registers).
// NLATE: the last NLATE bulk segments are issued separately (issue_late / wait_late, on the second mbarrier; single-
// buffered slabs only) for data whose shared-memory range is reused as scratch until the end of the step.
template <int G, int GR, int RL, int RQ, int NBULK = 0, int NLATE = 0>
struct SlabLoader {
  static constexpr int RQ1 = RQ > 0 ? RQ : 1;
  static constexpr int NB1 = NBULK > 0 ? NBULK : 1;
  static constexpr int NEARLY = NBULK - NLATE;
  CopyPlan<RL, GR> main;
  CopyPlan<RQ1, GR> sh;
  unsigned pidx[G];  // (clamped problem index) * horizon, per group (warp-uniform)
  int group_bytes, lane, shift;
  // bulk segments (warp-uniform)
  const char* bgp[NB1];
  int bstride[NB1], bbytes[NB1], bsoff[NB1], nbulk_added, btotal, btotal_late;
  uint64_t* bars;
  uint32_t phase;

  IDOC_DEVICE void init(long long prob0, long long last, int Th, int group_bytes_, int lane_, uint64_t* bars_ = nullptr) {
    main.reset();
    sh.reset();
    group_bytes = group_bytes_;
    lane = lane_;
    shift = 0;
    nbulk_added = btotal = btotal_late = 0;
    bars = bars_;
    phase = 0;
#pragma unroll
    for (int g = 0; g < G; g++) {
      long long p = prob0 + g;
      p = p <= last ? p : last;  // idle groups shadow the last problem (their stores are masked)
      pidx[g] = (unsigned)(p * Th);
    }
    if (NBULK > 0) {
      if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
      }
      __syncwarp();
    }
  }
  // a large segment moved by the TMA engine (call exactly NBULK times); sub_off, sub_bytes and smem_off multiples of 16
  IDOC_DEVICE void add_bulk(const void* base, int seg_bytes, int sub_off, int sub_bytes, int smem_off) {
    if (NBULK > 0) {
#pragma unroll
      for (int s = 0; s < NB1; s++)
        if (s == nbulk_added) {
          bgp[s] = static_cast<const char*>(base) + sub_off;
          bstride[s] = seg_bytes;
          bbytes[s] = sub_bytes;
          bsoff[s] = smem_off;
        }
      if (nbulk_added < NEARLY) btotal += G * sub_bytes; else btotal_late += G * sub_bytes;
      nbulk_added++;
    }
  }
  IDOC_DEVICE void add(const void* base, int seg_bytes, int sub_off, int sub_bytes, int smem_off, int tshift = 0) {
    if (tshift == 0) {
      main.add(base, seg_bytes, sub_off, sub_bytes, smem_off, lane);
    } else {
      sh.add(base, seg_bytes, sub_off, sub_bytes, smem_off, lane);
      shift = tshift;
    }
  }
  // groups_base32: 32-bit shared address of group 0's slab; buf_off: byte offset of the in-slab buffer;
  // buf: which mbarrier tracks this buffer (0/1).  The caller guarantees (by a __syncwarp) that every lane has
  // finished reading the buffer being refilled.  Must be called by all 32 lanes, converged.
  IDOC_DEVICE void issue(uint32_t groups_base32, int buf_off, int t, int buf = 0) const {
    if (NEARLY > 0) {
      if (elect_one()) {
        mbar_expect_tx(&bars[buf], (uint32_t)btotal);
#pragma unroll
        for (int g = 0; g < G; g++)
#pragma unroll
          for (int s = 0; s < NEARLY; s++)
            bulk_g2s(groups_base32 + g * group_bytes + buf_off + bsoff[s],
                     bgp[s] + (unsigned long long)(pidx[g] + (unsigned)t) * (unsigned)bstride[s], (uint32_t)bbytes[s], &bars[buf]);
      }
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
    if (NEARLY > 0) {
      mbar_wait(&bars[buf], (phase >> buf) & 1u);
      phase ^= (1u << buf);
    }
  }
  // the late segments of step t (all 32 lanes, converged; buffer 0 only)
  IDOC_DEVICE void issue_late(uint32_t groups_base32, int buf_off, int t) const {
    if (NLATE > 0) {
      if (elect_one()) {
        mbar_expect_tx(&bars[1], (uint32_t)btotal_late);
#pragma unroll
        for (int g = 0; g < G; g++)
#pragma unroll
          for (int s = NEARLY; s < NBULK; s++)
            bulk_g2s(groups_base32 + g * group_bytes + buf_off + bsoff[s],
                     bgp[s] + (unsigned long long)(pidx[g] + (unsigned)t) * (unsigned)bstride[s], (uint32_t)bbytes[s], &bars[1]);
      }
    }
  }
  IDOC_DEVICE void wait_late() {
    if (NLATE > 0) {
      mbar_wait(&bars[1], (phase >> 1) & 1u);
      phase ^= 2u;
    }
  }
};

// shared -> global by bulk async copies: NSB large segments per group, all issued by one elected lane from
// warp-uniform addresses.  Protocol per step: every lane that wrote the slab executes fence_proxy_async();
// __syncwarp(); store(); and wait_read() (+ __syncwarp()) before the slab is written again; drain() before the warp exits.
template <int G, int NSB>
struct BulkStorer {
  static constexpr int NS1 = NSB > 0 ? NSB : 1;
  char* gp[NS1];
  unsigned idx0;
  int stride[NS1], bytes[NS1], soff[NS1], tsh[NS1], group_bytes, nadded, ngroups;
  IDOC_DEVICE void init(long long prob0, long long last, int Th, int group_bytes_, int /*lane*/) {
    group_bytes = group_bytes_;
    idx0 = (unsigned)(prob0 * Th);
    const long long ng = last - prob0 + 1;
    ngroups = ng < G ? (int)ng : G;
    nadded = 0;
  }
  // tshift: the step-t store of this segment goes to timestep t + tshift (skipped when that is negative)
  IDOC_DEVICE void add(void* base, int seg_bytes, int sub_off, int sub_bytes, int smem_off, int tshift = 0) {
    if (NSB > 0) {
#pragma unroll
      for (int s = 0; s < NS1; s++)
        if (s == nadded) {
          gp[s] = static_cast<char*>(base) + sub_off;
          stride[s] = seg_bytes;
          bytes[s] = sub_bytes;
          soff[s] = smem_off;
          tsh[s] = tshift;
        }
      nadded++;
    }
  }
  // Th: horizon (slab index of group g = idx0 + g * Th + t).  All 32 lanes, converged.
  IDOC_DEVICE void store(uint32_t groups_base32, int buf_off, int t, int Th, bool shifted_only = false) const {
    if (NSB > 0) {
      if (elect_one()) {
#pragma unroll
        for (int s = 0; s < NS1; s++)
          if (t + tsh[s] >= 0 && !(shifted_only && tsh[s] == 0)) {
            char* base = gp[s] + (unsigned long long)(idx0 + (unsigned)(t + tsh[s])) * (unsigned)stride[s];
            const unsigned long long gstep = (unsigned long long)Th * (unsigned)stride[s];
            const uint32_t sbase = groups_base32 + buf_off + soff[s];
            if (ngroups == G) {
#pragma unroll
              for (int g = 0; g < G; g++) bulk_s2g(base + g * gstep, sbase + g * group_bytes, (uint32_t)bytes[s]);
            } else {
              for (int g = 0; g < ngroups; g++) bulk_s2g(base + g * gstep, sbase + g * group_bytes, (uint32_t)bytes[s]);
            }
          }
        bulk_commit();
      }
    }
  }
  // one copy per group of `len` bytes starting at byte `lo` of the group's buffer, to byte `goff` of slab t of segment 0
  // (the range may run over the end of slab t into slab t+1: slabs of consecutive timesteps are adjacent)
  IDOC_DEVICE void store_range(uint32_t groups_base32, int buf_off, int t, int Th, int goff, int lo, int len) const {
    if (NSB > 0) {
      if (elect_one()) {
        char* base = gp[0] + (unsigned long long)(idx0 + (unsigned)t) * (unsigned)stride[0] + goff;
        const unsigned long long gstep = (unsigned long long)Th * (unsigned)stride[0];
        const uint32_t sbase = groups_base32 + buf_off + lo;
        if (ngroups == G) {
#pragma unroll
          for (int g = 0; g < G; g++) bulk_s2g(base + g * gstep, sbase + g * group_bytes, (uint32_t)len);
        } else {
          for (int g = 0; g < ngroups; g++) bulk_s2g(base + g * gstep, sbase + g * group_bytes, (uint32_t)len);
        }
        bulk_commit();
      }
    }
  }
  IDOC_DEVICE void wait_read() const {
    if (NSB > 0) bulk_wait_read0();
  }
  IDOC_DEVICE void drain() const {
    if (NSB > 0) bulk_wait_all0();
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
