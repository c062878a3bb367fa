// Slab movers shared by the time-sweep kernels: one interface, two engines.
//   TMA == true  (every segment a multiple of 16 B): 1-D bulk async copies completing on per-warp mbarriers
//   TMA == false (odd shapes, e.g. n = 3 in fp32): per-lane cp.async of GR-byte chunks
#pragma once
#include "common.cuh"

namespace idoc {

constexpr int WARP_HDR_BYTES = 16;  // two mbarriers at the start of every warp's shared-memory region

template <bool TMA, int G, int GR, int NSEG, int RL, int RQ>
struct SlabLoader;

template <int G, int GR, int NSEG, int RL, int RQ>
struct SlabLoader<false, G, GR, NSEG, RL, RQ> {
  CopyPlan<RL, GR> main;
  CopyPlan<(RQ > 0 ? RQ : 1), GR> sh;
  long long prob0, last;
  int Th, group_bytes, lane, shift;

  IDOC_DEVICE void init(char* /*warp_region*/, long long prob0_, long long last_, int Th_, int group_bytes_, int lane_) {
    main.reset();
    sh.reset();
    prob0 = prob0_; last = last_; Th = Th_; group_bytes = group_bytes_; lane = lane_; shift = 0;
  }
  IDOC_DEVICE void add(const void* base, int seg_bytes, int sub_off, int sub_bytes, int smem_off, int tshift = 0) {
    if (tshift == 0) {
      main.add(base, seg_bytes, sub_off, sub_bytes, smem_off, lane);
    } else {
      sh.add(base, seg_bytes, sub_off, sub_bytes, smem_off, lane);
      shift = tshift;
    }
  }
  IDOC_DEVICE void issue(char* groups_base, int buf_off, int t, int /*buf*/) {
#pragma unroll
    for (int g = 0; g < G; g++) {
      long long p = prob0 + g;
      p = p <= last ? p : last;
      char* slab = groups_base + g * group_bytes + buf_off;
      main.load(slab, p * Th + t);
      if (RQ > 0 && t + shift >= 0) sh.load(slab, p * Th + t + shift);
    }
    cp_async_commit();
  }
  IDOC_DEVICE void wait(int /*buf*/) {
    cp_async_wait<0>();
    __syncwarp();
  }
};

template <int G, int GR, int NSEG, int RL, int RQ>
struct SlabLoader<true, G, GR, NSEG, RL, RQ> {
  BulkPlan<cdiv(NSEG * G, 32)> plan;
  uint64_t* bars;
  uint32_t phase;
  long long prob0, last;
  int Th, group_bytes, lane;

  IDOC_DEVICE void init(char* warp_region, long long prob0_, long long last_, int Th_, int group_bytes_, int lane_) {
    plan.reset();
    bars = reinterpret_cast<uint64_t*>(warp_region);
    phase = 0;
    prob0 = prob0_; last = last_; Th = Th_; group_bytes = group_bytes_; lane = lane_;
    if (lane == 0) {
      mbar_init(&bars[0], 1);
      mbar_init(&bars[1], 1);
      fence_mbar_init();
    }
    __syncwarp();
  }
  IDOC_DEVICE void add(const void* base, int seg_bytes, int sub_off, int sub_bytes, int smem_off, int tshift = 0) {
    plan.template add<G>(base, seg_bytes, sub_off, sub_bytes, smem_off, group_bytes, prob0, last, Th, lane, tshift, false);
  }
  IDOC_DEVICE void issue(char* groups_base, int buf_off, int t, int buf) { plan.load(groups_base + buf_off, t, &bars[buf], lane); }
  IDOC_DEVICE void wait(int buf) {
    mbar_wait(&bars[buf], (phase >> buf) & 1u);
    phase ^= (1u << buf);
  }
};

// ---------------------------------------------------------------------------------------------- stores
template <bool TMA, int G, int GR, int NSEG, int RS>
struct SlabStorer;

template <int G, int GR, int NSEG, int RS>
struct SlabStorer<false, G, GR, NSEG, RS> {
  CopyPlan<RS, GR> st;
  long long prob0, last;
  int Th, group_bytes, lane;
  IDOC_DEVICE void init(long long prob0_, long long last_, int Th_, int group_bytes_, int lane_) {
    st.reset();
    prob0 = prob0_; last = last_; Th = Th_; group_bytes = group_bytes_; lane = lane_;
  }
  IDOC_DEVICE void add(void* base, int seg_bytes, int sub_off, int sub_bytes, int smem_off) {
    st.add(base, seg_bytes, sub_off, sub_bytes, smem_off, lane);
  }
  // all lanes' shared-memory writes must be complete (a __syncwarp()) before calling
  IDOC_DEVICE void store(const char* groups_base, int buf_off, int t) {
#pragma unroll
    for (int g = 0; g < G; g++) {
      const long long p = prob0 + g;
      if (p <= last) st.store(groups_base + g * group_bytes + buf_off, p * Th + t);
    }
  }
  IDOC_DEVICE void drain() {}
};

template <int G, int GR, int NSEG, int RS>
struct SlabStorer<true, G, GR, NSEG, RS> {
  BulkPlan<cdiv(NSEG * G, 32)> plan;
  long long prob0, last;
  int Th, group_bytes, lane;
  IDOC_DEVICE void init(long long prob0_, long long last_, int Th_, int group_bytes_, int lane_) {
    plan.reset();
    prob0 = prob0_; last = last_; Th = Th_; group_bytes = group_bytes_; lane = lane_;
  }
  IDOC_DEVICE void add(void* base, int seg_bytes, int sub_off, int sub_bytes, int smem_off) {
    plan.template add<G>(base, seg_bytes, sub_off, sub_bytes, smem_off, group_bytes, prob0, last, Th, lane, 0, true);
  }
  // the generic-proxy writes of all lanes are fenced towards the async proxy, then the bulk stores are issued
  IDOC_DEVICE void store(const char* groups_base, int buf_off, int t) {
    fence_proxy_async();
    __syncwarp();
    plan.store(groups_base + buf_off, t);
  }
  // blocks until the bulk stores issued so far have finished READING shared memory
  IDOC_DEVICE void drain() {
    bulk_wait_read<0>();
    __syncwarp();
  }
};

}  // namespace idoc
