// Access-pattern roofline for the time-sweep kernels: moves exactly the bytes one sweep moves, with the same
// traversal (a warp owns 8 problems and walks t; per step and problem it reads one contiguous segment of each
// input array and writes one of each output array; consecutive steps are adjacent in memory), but does no
// arithmetic.  TB > 1 moves TB consecutive timesteps per request (segments TB times longer, T/TB iterations).
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/pattern_bench tools/pattern_bench.cu
//   tools/pattern_bench            (prints one line per configuration)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);   \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

constexpr int MAXSEG = 10;
struct Pattern {
  int nr, nw;                 // number of read / written arrays
  int rbytes[MAXSEG];         // bytes per (problem, request) of each read array
  int wbytes[MAXSEG];
  int rper[MAXSEG];           // a request every rper steps (moves rper consecutive timesteps)
  int wper[MAXSEG];
  char* rptr[MAXSEG];
  char* wptr[MAXSEG];
  int rchunks, wchunks;       // 16-byte chunks per (problem, step)
  int steps;                  // iterations (T / TB)
  long long nb;
  int nbuf;
};

__device__ __forceinline__ void cp16(unsigned dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
}

// one warp = 8 problems; smem per warp = nbuf * 8 * rchunks * 16 bytes.  Flat chunk numbering over the segments of
// a problem-step (as SlabLoader does): lane handles chunks lane, lane+32, ... (R rounds) of each of its 8 problems.
template <int RR, int RW>
__global__ void pattern_kernel(const Pattern p) {
  extern __shared__ __align__(16) char sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const long long prob0 = ((long long)blockIdx.x * wpb + warp) * 8;
  if (prob0 >= p.nb) return;
  const int per_buf = 8 * p.rchunks * 16;
  char* wb = sm + (size_t)warp * p.nbuf * per_buf;
  const unsigned wb32 = (unsigned)__cvta_generic_to_shared(wb);
  const char* rp[RR];
  int rs[RR], rpe[RR], wpe[RW > 0 ? RW : 1];
  char* wp[RW > 0 ? RW : 1];
  int ws[RW > 0 ? RW : 1];
#pragma unroll
  for (int r = 0; r < RR; r++) {
    int c = lane + 32 * r, c0 = 0;
    rp[r] = nullptr;
    rs[r] = 0;
    for (int s = 0; s < p.nr; s++) {
      const int nc = p.rbytes[s] >> 4;
      if (c >= c0 && c < c0 + nc) {
        rp[r] = p.rptr[s] + (c - c0) * 16;
        rs[r] = p.rbytes[s];
        rpe[r] = p.rper[s] == 8 ? 3 : p.rper[s] == 4 ? 2 : p.rper[s] == 2 ? 1 : 0;
      }
      c0 += nc;
    }
  }
#pragma unroll
  for (int r = 0; r < RW; r++) {
    int c = lane + 32 * r, c0 = 0;
    wp[r] = nullptr;
    ws[r] = 0;
    for (int s = 0; s < p.nw; s++) {
      const int nc = p.wbytes[s] >> 4;
      if (c >= c0 && c < c0 + nc) {
        wp[r] = p.wptr[s] + (c - c0) * 16;
        ws[r] = p.wbytes[s];
        wpe[r] = p.wper[s] == 8 ? 3 : p.wper[s] == 4 ? 2 : p.wper[s] == 2 ? 1 : 0;
      }
      c0 += nc;
    }
  }
  unsigned pidx[8];
#pragma unroll
  for (int g = 0; g < 8; g++) pidx[g] = (unsigned)((prob0 + g < p.nb ? prob0 + g : p.nb - 1) * p.steps);

  auto issue = [&](int it, int buf) {
#pragma unroll
    for (int g = 0; g < 8; g++) {
#pragma unroll
      for (int r = 0; r < RR; r++)
        if (rs[r] && (it & ((1 << rpe[r]) - 1)) == 0)
          cp16(wb32 + buf * per_buf + (g * p.rchunks + lane + 32 * r) * 16, rp[r] + (unsigned long long)((pidx[g] + it) >> rpe[r]) * (unsigned)rs[r]);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  issue(0, 0);
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (int it = 0; it < p.steps; it++) {
    const int buf = it % p.nbuf;
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncwarp();
    if (p.nbuf > 1 && it + 1 < p.steps) issue(it + 1, (it + 1) % p.nbuf);
    const uint4 v = *reinterpret_cast<const uint4*>(wb + buf * per_buf + lane * 16);
    acc.x ^= v.x; acc.y += v.y; acc.z ^= v.z; acc.w += v.w;
#pragma unroll
    for (int g = 0; g < 8; g++) {
      if (prob0 + g < p.nb) {
#pragma unroll
        for (int r = 0; r < RW; r++)
          if (ws[r] && (it & ((1 << wpe[r]) - 1)) == (1 << wpe[r]) - 1)
            __stcs(reinterpret_cast<uint4*>(wp[r] + (unsigned long long)((pidx[g] + it) >> wpe[r]) * (unsigned)ws[r]), acc);
      }
    }
    __syncwarp();
    if (p.nbuf == 1 && it + 1 < p.steps) issue(it + 1, 0);
  }
}

typedef void (*KernT)(const Pattern);
template <int RR>
static KernT pick_w(int rw) {
  switch (rw) {
    case 0: return pattern_kernel<RR, 0>;
    case 1: return pattern_kernel<RR, 1>;
    case 2: return pattern_kernel<RR, 2>;
    case 3: return pattern_kernel<RR, 3>;
    case 4: return pattern_kernel<RR, 4>;
    case 6: case 5: return pattern_kernel<RR, 6>;
    default: return pattern_kernel<RR, 8>;
  }
}
static KernT pick(int rr, int rw) {
  switch (rr) {
    case 1: return pick_w<1>(rw);
    case 2: return pick_w<2>(rw);
    case 3: return pick_w<3>(rw);
    case 4: return pick_w<4>(rw);
    case 5: case 6: return pick_w<6>(rw);
    default: return pick_w<8>(rw);
  }
}

// linear streaming kernels (grid-stride over 16-byte chunks): write-only, read-only, and read r / write w chunk mixes
__global__ void lin_write(uint4* dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint4 v = make_uint4(1, 2, 3, 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) __stcs(dst + i, v);
}
__global__ void lin_read(const uint4* src, size_t n, uint4* sink) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint4 v = __ldcs(src + i);
    acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
  }
  if (acc.x == 0x12345678u) *sink = acc;
}
// every thread: reads 1 chunk, writes W chunks (W = 1: copy)
template <int W>
__global__ void lin_mix(const uint4* src, uint4* dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint4 v = __ldcs(src + i);
#pragma unroll
    for (int w = 0; w < W; w++) __stcs(dst + (size_t)w * n + i, v);
  }
}

struct Cfg {
  const char* name;
  std::vector<int> r, w;    // bytes per step of each array
  std::vector<int> rp, wp;  // request period (steps) of each array
};

int main(int argc, char** argv) {
  const long long B = argc > 1 ? atoll(argv[1]) : 65536;
  const int T = 96;
  std::vector<Cfg> cfgs = {
      {"adj_fwd f32 base", {256, 128, 128, 144, 48, 16, 32, 32}, {256, 128, 256, 128, 64, 32, 16, 32}, {1, 1, 1, 1, 1, 1, 1, 1}, {1, 1, 1, 1, 1, 1, 1, 1}},
      {"adj_fwd f32 small w x4", {256, 128, 128, 144, 48, 16, 32, 32}, {256, 128, 256, 128, 64, 32, 16, 32}, {1, 1, 1, 1, 1, 1, 1, 1}, {1, 1, 1, 1, 4, 4, 4, 4}},
      {"adj_fwd f32 big w x2 small w x4", {256, 128, 128, 144, 48, 16, 32, 32}, {256, 128, 256, 128, 64, 32, 16, 32}, {1, 1, 1, 1, 1, 1, 1, 1}, {2, 2, 2, 2, 4, 4, 4, 4}},
      {"adj_fwd f32 all w x4", {256, 128, 128, 144, 48, 16, 32, 32}, {256, 128, 256, 128, 64, 32, 16, 32}, {1, 1, 1, 1, 1, 1, 1, 1}, {4, 4, 4, 4, 4, 4, 4, 4}},
      {"adj_fwd f32 r x2, w x2/x4", {256, 128, 128, 144, 48, 16, 32, 32}, {256, 128, 256, 128, 64, 32, 16, 32}, {2, 2, 2, 2, 4, 4, 4, 4}, {2, 2, 2, 2, 4, 4, 4, 4}},
      {"adj_fwd f64 small w x4", {512, 256, 256, 288, 96, 32, 64, 64}, {512, 256, 512, 256, 128, 64, 32, 64}, {1, 1, 1, 1, 1, 1, 1, 1}, {1, 1, 1, 1, 4, 4, 4, 4}},
      {"adj_fwd f64 big w x2 small w x4", {512, 256, 256, 288, 96, 32, 64, 64}, {512, 256, 512, 256, 128, 64, 32, 64}, {1, 1, 1, 1, 1, 1, 1, 1}, {2, 2, 2, 2, 4, 4, 4, 4}},
      {"rollout f32 base", {256, 128, 32, 320}, {32, 16, 32}, {1, 1, 1, 1}, {4, 4, 4}},
      {"rollout f32 r x2", {256, 128, 32, 320}, {32, 16, 32}, {2, 2, 2, 2}, {4, 4, 4}},
      {"adj_bwd f32 base", {256, 128, 176, 16, 32}, {48}, {1, 1, 1, 1, 1}, {1}},
      {"adj_bwd f32 r x2", {256, 128, 176, 16, 32}, {48}, {2, 2, 2, 2, 2}, {2}},
      {"riccati f32 base", {256, 128, 256, 128, 64, 32, 16, 32}, {368}, {1, 1, 1, 1, 1, 1, 1, 1}, {1}},
      {"riccati f32 r x2 w x2", {256, 128, 256, 128, 64, 32, 16, 32}, {368}, {2, 2, 2, 2, 2, 2, 2, 2}, {2}},
  };
  size_t maxbytes = 0;
  for (auto& c : cfgs) {
    size_t s = 0;
    for (int x : c.r) s += x;
    for (int x : c.w) s += x;
    if (s * B * T > maxbytes) maxbytes = s * B * T;
  }
  char* arena;
  CK(cudaMalloc(&arena, maxbytes + (1 << 20)));
  CK(cudaMemset(arena, 1, maxbytes));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  printf("B=%lld T=%d\n", B, T);
  {
    const size_t n = (size_t)4 << 30 >> 4;  // 4 GiB of chunks
    auto timeit = [&](const char* name, double bytes, auto launch) {
      launch();
      CK(cudaDeviceSynchronize());
      float best = 1e30f;
      for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
      }
      printf("%-28s %7.3f ms  %7.1f GB/s\n", name, best, bytes / (best * 1e-3) / 1e9);
    };
    uint4* a4 = reinterpret_cast<uint4*>(arena);
    for (int blocks : {148 * 8, 148 * 16, 148 * 32}) {
      printf("linear, %d blocks x 256 threads\n", blocks);
      timeit("  write only 4 GiB", 16.0 * n, [&] { lin_write<<<blocks, 256>>>(a4, n); });
      timeit("  read only 4 GiB", 16.0 * n, [&] { lin_read<<<blocks, 256>>>(a4, n, a4 + 2 * n); });
      timeit("  copy 2+2 GiB", 16.0 * n, [&] { lin_mix<1><<<blocks, 256>>>(a4, a4 + n / 2, n / 2); });
      timeit("  read 1 : write 5 (0.5 + 2.5 GiB)", 16.0 * (n / 8) * 6, [&] { lin_mix<5><<<blocks, 256>>>(a4, a4 + n / 8, n / 8); });
    }
  }
  for (auto& c : cfgs) {
    for (int warps_per_sm : {4, 6, 8, 12}) {
      const int nbuf = 2;
      Pattern p;
      memset(&p, 0, sizeof(p));
      p.nr = (int)c.r.size();
      p.nw = (int)c.w.size();
      p.steps = T;
      p.nb = B;
      p.nbuf = nbuf;
      size_t off = 0, tot = 0;
      for (int i = 0; i < p.nr; i++) {
        p.rper[i] = c.rp[i];
        p.rbytes[i] = c.r[i] * c.rp[i];
        p.rptr[i] = arena + off;
        off += (size_t)c.r[i] * B * T;
        p.rchunks += p.rbytes[i] / 16;
        tot += c.r[i];
      }
      for (int i = 0; i < p.nw; i++) {
        p.wper[i] = c.wp[i];
        p.wbytes[i] = c.w[i] * c.wp[i];
        p.wptr[i] = arena + off;
        off += (size_t)c.w[i] * B * T;
        p.wchunks += p.wbytes[i] / 16;
        tot += c.w[i];
      }
      const size_t warp_smem = (size_t)nbuf * 8 * p.rchunks * 16;
      size_t want = (size_t)(226 * 1024) / warps_per_sm - 1024;  // per-warp budget
      if (warp_smem > want) continue;
      size_t smem = want > 227 * 1024 ? 227 * 1024 : want;
      const int rr = (p.rchunks + 31) / 32, rw = (p.wchunks + 31) / 32;
      if (rr > 8 || rw > 8) { printf("%-28s: too many rounds (%d, %d), skipped\n", c.name, rr, rw); continue; }
      KernT pattern_kernel = pick(rr, rw);
      CK(cudaFuncSetAttribute(pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const unsigned grid = (unsigned)((B + 7) / 8);
      pattern_kernel<<<grid, 32, smem>>>(p);
      CK(cudaDeviceSynchronize());
      float best = 1e30f;
      for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        pattern_kernel<<<grid, 32, smem>>>(p);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
      }
      printf("%-28s warps/SM=%2d  %7.3f ms  %7.1f GB/s\n", c.name, warps_per_sm, best, (double)tot * B * T / (best * 1e-3) / 1e9);
      fflush(stdout);
    }
  }
  return 0;
}
