// ssd_kernels_points.cuh -- the per-point (HBM-bound) kernels of the chain:
//   k_transform_bin : z>0 filter, camera->world transform, range filter, height bin, height histogram
//   k_peaks         : histogram peaks -> plateau bands -> bin->label LUT, ground selection
//   k_label_bev     : per-point segment label + top-down (BEV) occupancy bitmaps of the outlined plateaus
//   k_quad_reduce   : point-in-quadrilateral filter, per-step z sum / count, ground BEV bitmap
// Algorithmic traffic: 12 B read + 1 B written per point; the bin codes are written once by
// k_transform_bin and rewritten in place as labels by k_label_bev.
#pragma once
#include "ssd_device.cuh"

#define SSD_PT_THREADS 256
#define SSD_PT_WARPS (SSD_PT_THREADS / 32)

// 4 consecutive packed {x,y,z} vertices = 3 float4
struct Quad4
{
  float x[4], y[4], z[4];
};

__device__ __forceinline__ float4 ldg_stream(const float4 *p)
{
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ Quad4 load_quad(const float4 *__restrict__ xyz4, size_t q)
{
  const float4 v0 = __ldg(xyz4 + q * 3), v1 = __ldg(xyz4 + q * 3 + 1), v2 = __ldg(xyz4 + q * 3 + 2);
  Quad4 r;
  r.x[0] = v0.x; r.y[0] = v0.y; r.z[0] = v0.z;
  r.x[1] = v0.w; r.y[1] = v1.x; r.z[1] = v1.y;
  r.x[2] = v1.z; r.y[2] = v1.w; r.z[2] = v2.x;
  r.x[3] = v2.y; r.y[3] = v2.z; r.z[3] = v2.w;
  return r;
}

// ---------------------------------------------------------------------------------------------
// k_transform_bin: PointsExtraction::extract + HeightsHistogram::calcHist
// (pointcloud.cpp:122-178, 194-204). grid = (tiles_per_frame, frames), block = 256.
// Each thread handles 4 consecutive points per iteration: three 16 B loads, one 4 B store.
// The camera->world transform, range filter and height bin are decided in single precision with a rigorous
// error bound (point_code_filtered); the exact double-precision chain runs only for the few points whose
// f32 value lies within that bound of a threshold, so the result is bit-identical to the all-double chain.
// Histogram: each thread run-length merges its own codes (neighbouring pixels mostly share a bin) and adds
// the runs to one shared-memory histogram per block; one global atomic per non-empty bin per block.
// ---------------------------------------------------------------------------------------------
template<int ITERS>
__global__ void __launch_bounds__(SSD_PT_THREADS) k_transform_bin(const __grid_constant__ DevParams p, const float *__restrict__ xyz,
                                                                   unsigned char *__restrict__ codes, FrameDev *__restrict__ frames,
                                                                   unsigned long long *__restrict__ n_exact)
{
  __shared__ unsigned s_hist[SSD_BINS_PAD];
  const int tid = threadIdx.x;
  const int frame = blockIdx.y;
  s_hist[tid] = 0; // SSD_PT_THREADS == SSD_BINS_PAD
  __syncthreads();

  const size_t fbase = (size_t)frame * p.N;
  const float4 *xyz4 = reinterpret_cast<const float4 *>(xyz + fbase * 3);
  unsigned *codes32 = reinterpret_cast<unsigned *>(codes + fbase);
  const int nquads = p.N >> 2;
  unsigned run_code = SSD_CODE_INVALID, run_n = 0, exact = 0; // an empty run of a valid code: no sentinel pattern to collide with

#pragma unroll
  for(int it = 0; it < ITERS; it++)
  {
    const int q = (blockIdx.x * ITERS + it) * SSD_PT_THREADS + tid;
    if(q < nquads)
    {
      const Quad4 v = load_quad(xyz4, q);
      unsigned c[4];
      bool unc[4];
#pragma unroll
      for(int j = 0; j < 4; j++)
        c[j] = point_code_filtered(p, v.x[j], v.y[j], v.z[j], unc[j]);
      unsigned um = (unc[0] ? 1u : 0u) | (unc[1] ? 2u : 0u) | (unc[2] ? 4u : 0u) | (unc[3] ? 8u : 0u);
      if(um)
      {
        // rare: one out-of-line exact evaluation per uncertain point
        exact += __popc(um);
#pragma unroll 1
        while(um)
        {
          const int j = __ffs(um) - 1;
          um &= um - 1;
          const float fx = j == 0 ? v.x[0] : (j == 1 ? v.x[1] : (j == 2 ? v.x[2] : v.x[3]));
          const float fy = j == 0 ? v.y[0] : (j == 1 ? v.y[1] : (j == 2 ? v.y[2] : v.y[3]));
          const float fz = j == 0 ? v.z[0] : (j == 1 ? v.z[1] : (j == 2 ? v.z[2] : v.z[3]));
          const unsigned ce = point_code_slow(p, fx, fy, fz);
          c[0] = j == 0 ? ce : c[0];
          c[1] = j == 1 ? ce : c[1];
          c[2] = j == 2 ? ce : c[2];
          c[3] = j == 3 ? ce : c[3];
        }
      }
      const unsigned cw = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
      codes32[q] = cw;
      if(cw == run_code * 0x01010101u)
        run_n += 4; // all four in the current run (neighbouring pixels mostly share a bin)
      else
      {
#pragma unroll
        for(int j = 0; j < 4; j++)
        {
          if(c[j] == run_code)
            run_n++;
          else
          {
            if(run_n)
              atomicAdd(&s_hist[run_code], run_n);
            run_code = c[j];
            run_n = 1;
          }
        }
      }
    }
  }
  if(run_n)
    atomicAdd(&s_hist[run_code], run_n);
  if(exact && n_exact)
    atomicAdd(n_exact, (unsigned long long)exact);
  __syncthreads();
  const unsigned sum = s_hist[tid];
  if(sum)
    atomicAdd(&frames[frame].hist[tid], sum);
}

// ---------------------------------------------------------------------------------------------
// k_peaks: HeightsHistogram::findPeaks/filterPeaks (pointcloud.cpp:214-256), the plateau bands of
// PlateausExtraction::extractPlateauPoints (:300-335) folded into a bin->label LUT, and the ground /
// first-outlined bookkeeping of StairsDetector::detectStairSteps (:402-418).
// One warp per frame: the histogram and the LUT live in shared memory, lane 0 walks the <= 253 bins,
// the lanes write the LUT and the plateau records back in parallel.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_peaks(const __grid_constant__ DevParams p, FrameDev *__restrict__ frames, int n_frames)
{
  __shared__ unsigned s_hist[SSD_BINS_PAD];
  __shared__ __align__(8) unsigned char s_lut[SSD_BINS_PAD];
  __shared__ int s_height[SSD_GPU_MAX_PLATEAUS], s_hmin[SSD_GPU_MAX_PLATEAUS], s_hmax[SSD_GPU_MAX_PLATEAUS];
  __shared__ unsigned s_np[SSD_GPU_MAX_PLATEAUS];
  __shared__ int s_K, s_first_outlined;
  const int f = blockIdx.x, lane = threadIdx.x;
  if(f >= n_frames)
    return;
  FrameDev &F = frames[f];
  for(int b = lane; b < SSD_BINS_PAD; b += 32)
  {
    s_hist[b] = F.hist[b];
    s_lut[b] = (unsigned char)SSD_LABEL_REMAINDER;
  }
  __syncwarp();
  if(lane == 0)
  {
    const unsigned *hist = s_hist;
    unsigned status = 0;
    s_lut[SSD_CODE_OUT_OF_RANGE] = (unsigned char)SSD_LABEL_OUT_OF_RANGE;
    s_lut[SSD_CODE_INVALID] = (unsigned char)SSD_LABEL_INVALID;
    int K = 0;
    bool ascending = false, wrapped = false;
    const int last = p.n_bins - 1;
    for(int i = 0; i < last; i++)
    {
      const unsigned c = hist[i], s = hist[i + 1];
      if(c < s)
      {
        ascending = true;
        continue;
      }
      if(c > s)
      {
        if(ascending && !(c < p.min_peak_points) && (unsigned)((c * 2u - hist[i - 1] - hist[i + 1]) * 2u) > c)
        {
          if(K >= SSD_GPU_MAX_PLATEAUS)
            status |= SSD_STATUS_TOO_MANY_PLATEAUS;
          else
          {
            int hmin, hmax;
            if(hist[i - 1] > hist[i + 1]) // :307-316
            {
              hmin = i - 1;
              hmax = i;
            }
            else
            {
              hmin = i;
              hmax = i + 1;
            }
            unsigned np = 0;
            if(hmin == 0)
            {
              // uint16 wrap of heightMin - 1 (:324): everything left goes to the remainder, this plateau
              // and all later ones stay empty
              wrapped = true;
              status |= SSD_STATUS_HMIN_WRAP;
            }
            if(!wrapped)
              for(int b = hmin; b <= hmax; b++)
                if(s_lut[b] == SSD_LABEL_REMAINDER)
                {
                  s_lut[b] = (unsigned char)K;
                  np += hist[b];
                }
            s_height[K] = i;
            s_hmin[K] = hmin;
            s_hmax[K] = hmax;
            s_np[K] = np;
            K++;
          }
        }
        ascending = false;
      }
    }
    int ground = -1, i = 0;
    unsigned maxGround = 0;
    for(; i < K; i++)
    {
      if(s_height[i] >= p.min_height)
        break;
      if(maxGround < s_np[i])
      {
        maxGround = s_np[i];
        ground = i;
      }
    }
    s_K = K;
    s_first_outlined = i;
    F.n_nonzero = (unsigned)p.N - hist[SSD_CODE_INVALID];
    F.n_in_range = (unsigned)p.N - hist[SSD_CODE_INVALID] - hist[SSD_CODE_OUT_OF_RANGE];
    F.n_plateaus = K;
    F.ground_index = ground;
    F.first_outlined = i;
    F.first_valid = -1;
    F.n_steps = 0;
    F.status = status;
  }
  __syncwarp();
  reinterpret_cast<unsigned long long *>(F.lut)[lane] = reinterpret_cast<const unsigned long long *>(s_lut)[lane];
  if(lane < s_K)
  {
    PlateauDev &P = F.plat[lane];
    P.height = s_height[lane];
    P.hmin = s_hmin[lane];
    P.hmax = s_hmax[lane];
    P.n_points = s_np[lane];
    P.valid = 0;
    P.outlined = lane >= s_first_outlined;
    P.n_in_quad = 0;
    P.quad_status = -1;
    P.mean_z = 0;
    P.sum_fix = 0;
    P.row_min = 0x7fffffff;
    P.row_max = -1;
    P.front_valid = 0;
    for(int c4 = 0; c4 < 4; c4++)
      P.quad_px[c4][0] = P.quad_px[c4][1] = P.quad_world[c4][0] = P.quad_world[c4][1] = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// k_label_bev: the per-point segment label (PlateausExtraction::extractPlateaus, pointcloud.cpp:280-343,
// as a LUT lookup) and StairsDetector::projectToBinaryImage (:458-471) for every outlined plateau.
// Rewrites the bin codes in place as labels. grid = (ceil(N / (1024*ITERS)), frames).
// Only the points of outlined plateaus (~1/4 of a frame) re-read their vertex and need the x,y rows of the
// transform; their BEV bits go to the plateau's global bitmap with atomicOr, the touched row range is
// tracked per thread run and merged through shared memory.
// ---------------------------------------------------------------------------------------------
template<int ITERS>
__global__ void __launch_bounds__(SSD_PT_THREADS) k_label_bev(const __grid_constant__ DevParams p, const float *__restrict__ xyz,
                                                               unsigned char *__restrict__ labels, FrameDev *__restrict__ frames,
                                                               unsigned *__restrict__ bev, size_t bm_words)
{
  __shared__ unsigned char s_lut[SSD_BINS_PAD];
  __shared__ int s_rmin[SSD_GPU_MAX_PLATEAUS], s_rmax[SSD_GPU_MAX_PLATEAUS];
  __shared__ unsigned s_oob;
  const int tid = threadIdx.x;
  const int frame = blockIdx.y;
  FrameDev &F = frames[frame];
  s_lut[tid] = F.lut[tid]; // SSD_PT_THREADS == SSD_BINS_PAD
  if(tid < SSD_GPU_MAX_PLATEAUS)
  {
    s_rmin[tid] = 0x7fffffff;
    s_rmax[tid] = -1;
  }
  if(tid == 0)
    s_oob = 0;
  const int first_outlined = F.first_outlined, K = F.n_plateaus;
  __syncthreads();

  const size_t fbase = (size_t)frame * p.N;
  const float4 *xyz4 = reinterpret_cast<const float4 *>(xyz + fbase * 3);
  unsigned *lab32 = reinterpret_cast<unsigned *>(labels + fbase);
  unsigned *fbev = bev + (size_t)frame * SSD_GPU_MAX_PLATEAUS * bm_words;
  const int nquads = p.N >> 2;
  unsigned oob = 0;
  int run_k = -1, run_min = 0x7fffffff, run_max = -1;
  // pending OR into one bitmap word
  unsigned *pend_addr = nullptr;
  unsigned pend_bits = 0;

#pragma unroll 2
  for(int it = 0; it < ITERS; it++)
  {
    const int q = (blockIdx.x * ITERS + it) * SSD_PT_THREADS + tid;
    if(q >= nquads)
      continue;
    const unsigned cw = lab32[q];
    unsigned l[4];
    bool any = false;
#pragma unroll
    for(int j = 0; j < 4; j++)
    {
      l[j] = s_lut[(cw >> (8 * j)) & 0xff];
      any |= (int)l[j] >= first_outlined && (int)l[j] < K;
    }
    lab32[q] = l[0] | (l[1] << 8) | (l[2] << 16) | (l[3] << 24);
    if(any)
    {
      const Quad4 v = load_quad(xyz4, q);
#pragma unroll
      for(int j = 0; j < 4; j++)
        if((int)l[j] >= first_outlined && (int)l[j] < K)
        {
          double wx, wy;
          camera_to_world_xy(p, v.x[j], v.y[j], v.z[j], wx, wy);
          int x, y;
          if(!bev_pixel(p, wx, wy, x, y))
          {
            oob = 1;
            continue;
          }
          if((int)l[j] != run_k)
          {
            if(run_max >= 0)
            {
              atomicMin(&s_rmin[run_k], run_min);
              atomicMax(&s_rmax[run_k], run_max);
            }
            run_k = (int)l[j];
            run_min = 0x7fffffff;
            run_max = -1;
          }
          run_min = min(run_min, y);
          run_max = max(run_max, y);
          unsigned *addr = fbev + (size_t)l[j] * bm_words + (size_t)y * p.wpr + (x >> 5);
          const unsigned bit = 1u << (x & 31);
          if(addr != pend_addr)
          {
            if(pend_bits)
              atomicOr(pend_addr, pend_bits);
            pend_addr = addr;
            pend_bits = 0;
          }
          pend_bits |= bit;
        }
    }
  }
  if(pend_bits)
    atomicOr(pend_addr, pend_bits);
  if(run_max >= 0)
  {
    atomicMin(&s_rmin[run_k], run_min);
    atomicMax(&s_rmax[run_k], run_max);
  }
  if(oob)
    s_oob = 1;
  __syncthreads();
  if(tid < K && s_rmax[tid] >= 0)
  {
    atomicMin(&F.plat[tid].row_min, s_rmin[tid]);
    atomicMax(&F.plat[tid].row_max, s_rmax[tid]);
  }
  if(tid == 0 && s_oob)
    atomicOr(&F.status, SSD_STATUS_BEV_OOB);
}

__device__ __forceinline__ long long warp_sum_s64(long long v)
{
#pragma unroll
  for(int s = 16; s > 0; s >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

// ---------------------------------------------------------------------------------------------
// k_quad_reduce: StairsDetector::getPointsInQuadrilateral + calcAverageZ (pointcloud.cpp:560-581) for the
// ground and every valid plateau, and the ground's BEV image (calcGround, :530-531).
// z is accumulated in 2^-36 m fixed point: integer sums are order independent, so the result is
// deterministic; the error (<= 2^-37 m per point) is eight orders below the 0.1 mm tolerance.
// Each thread sums its own run of points, the warp merges equal labels with shuffles (single-pass segmented
// reduce over the warp) and one lane issues the global 64-bit add.
// ---------------------------------------------------------------------------------------------
template<int ITERS>
__global__ void __launch_bounds__(SSD_PT_THREADS) k_quad_reduce(const __grid_constant__ DevParams p, const float *__restrict__ xyz,
                                                                 const unsigned char *__restrict__ labels, FrameDev *__restrict__ frames,
                                                                 unsigned *__restrict__ bev, size_t bm_words)
{
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ unsigned char s_active[SSD_BINS_PAD]; // label -> tested?
  __shared__ float4 s_box[SSD_GPU_MAX_PLATEAUS];   // verified inner box of each step (cx, hx, cy, hy)
  __shared__ int s_rmin, s_rmax;
  __shared__ unsigned s_oob;
  QuadTestDev *s_qt = reinterpret_cast<QuadTestDev *>(s_raw);
  const int tid = threadIdx.x;
  const int frame = blockIdx.y;
  FrameDev &F = frames[frame];
  if(F.first_valid < 0)
    return; // no valid plateau: nothing is emitted (pointcloud.cpp:434)
  const int K = F.n_plateaus, ground = F.ground_index;
  s_active[tid] = tid < K && F.plat[tid].valid && F.plat[tid].quad_status == 0; // SSD_PT_THREADS == SSD_BINS_PAD
  if(tid == 0)
  {
    s_rmin = 0x7fffffff;
    s_rmax = -1;
    s_oob = 0;
  }
  {
    const int words = (int)(sizeof(QuadTestDev) / 4);
    for(int i = tid; i < K * words; i += SSD_PT_THREADS)
    {
      const int k = i / words, w = i - k * words;
      reinterpret_cast<unsigned *>(&s_qt[k])[w] = reinterpret_cast<const unsigned *>(&F.plat[k].qt)[w];
    }
  }
  if(tid < K)
  {
    const QuadTestDev &g = F.plat[tid].qt;
    // the ground also needs the exact x,y of every accepted point for its BEV image: no fast accept there
    s_box[tid] = make_float4(g.ib_cx, tid == ground ? -1.f : g.ib_hx, g.ib_cy, g.ib_hy);
  }
  __syncthreads();

  const size_t fbase = (size_t)frame * p.N;
  const float4 *xyz4 = reinterpret_cast<const float4 *>(xyz + fbase * 3);
  const unsigned *lab32 = reinterpret_cast<const unsigned *>(labels + fbase);
  unsigned *gbev = ground >= 0 ? bev + ((size_t)frame * SSD_GPU_MAX_PLATEAUS + ground) * bm_words : nullptr;
  const int nquads = p.N >> 2;
  unsigned oob = 0;
  int rmin = 0x7fffffff, rmax = -1;

  long long acc = 0;
  unsigned acc_n = 0;
  int acc_k = -1;
#pragma unroll 2
  for(int it = 0; it < ITERS; it++)
  {
    const int q = (blockIdx.x * ITERS + it) * SSD_PT_THREADS + tid;
    if(q >= nquads)
      continue;
    const unsigned lw = __ldg(lab32 + q);
    const unsigned l0 = lw & 0xff, l1 = (lw >> 8) & 0xff, l2 = (lw >> 16) & 0xff, l3 = lw >> 24;
    if(!(s_active[l0] | s_active[l1] | s_active[l2] | s_active[l3]))
      continue;
    const Quad4 v = load_quad(xyz4, q);
#pragma unroll
    for(int j = 0; j < 4; j++)
    {
      const unsigned l = (lw >> (8 * j)) & 0xff;
      if(!s_active[l])
        continue;
      // fast accept: single-precision position inside the verified inner box by more than its error bound
      // (same bound as point_code_filtered: |w^ - w_ref| <= eps = E1 * max|p| + E0)
      const float4 bx = s_box[l];
      const float fx = v.x[j], fy = v.y[j], fz = v.z[j];
      const float eps = fmaf(p.E1, fmaxf(fmaxf(fabsf(fx), fabsf(fy)), fabsf(fz)), p.E0);
      const float wxf = fmaf(p.af[2], fz, fmaf(p.af[1], fy, fmaf(p.af[0], fx, p.bf[0])));
      const float wyf = fmaf(p.af[5], fz, fmaf(p.af[4], fy, fmaf(p.af[3], fx, p.bf[1])));
      const bool fast = fmaxf(fabsf(wxf - bx.x) - bx.y, fabsf(wyf - bx.z) - bx.w) < -eps;
      double wx = 0, wy = 0;
      if(!fast)
      {
        camera_to_world_xy(p, fx, fy, fz, wx, wy);
        if(!quadtest_within(s_qt[l], wx, wy))
          continue;
      }
      if((int)l != acc_k)
      {
        if(acc_n)
        {
          // label changed inside this thread's run (plateau border): flush straight to global memory
          atomicAdd(&F.plat[acc_k].sum_fix, (unsigned long long)acc);
          atomicAdd(&F.plat[acc_k].n_in_quad, acc_n);
        }
        acc = 0;
        acc_n = 0;
        acc_k = (int)l;
      }
      acc += z_to_fix(camera_to_world_z(p, v.x[j], v.y[j], v.z[j]));
      acc_n++;
      if((int)l == ground)
      {
        int x, y;
        if(!bev_pixel(p, wx, wy, x, y))
          oob = 1;
        else
        {
          atomicOr(gbev + (size_t)y * p.wpr + (x >> 5), 1u << (x & 31));
          rmin = min(rmin, y);
          rmax = max(rmax, y);
        }
      }
    }
  }
  // segmented reduce over the warp: one round per distinct label present
  {
    const int lane = tid & 31;
    unsigned todo = __ballot_sync(0xffffffffu, acc_n > 0);
    while(todo)
    {
      const int leader = __ffs(todo) - 1;
      const int kk = __shfl_sync(0xffffffffu, acc_k, leader);
      const bool mine = acc_n > 0 && acc_k == kk;
      const long long vs = warp_sum_s64(mine ? acc : 0ll);
      const unsigned ns = __reduce_add_sync(0xffffffffu, mine ? acc_n : 0u);
      if(lane == leader)
      {
        atomicAdd(&F.plat[kk].sum_fix, (unsigned long long)vs);
        atomicAdd(&F.plat[kk].n_in_quad, ns);
      }
      todo &= ~__ballot_sync(0xffffffffu, mine);
    }
    rmin = __reduce_min_sync(0xffffffffu, rmin);
    rmax = __reduce_max_sync(0xffffffffu, rmax);
    oob = __reduce_or_sync(0xffffffffu, oob);
    if(lane == 0)
    {
      if(rmax >= 0)
      {
        atomicMin(&s_rmin, rmin);
        atomicMax(&s_rmax, rmax);
      }
      if(oob)
        s_oob = 1;
    }
  }
  __syncthreads();
  if(tid == 0)
  {
    if(s_rmax >= 0)
    {
      atomicMin(&F.plat[ground].row_min, s_rmin);
      atomicMax(&F.plat[ground].row_max, s_rmax);
    }
    if(s_oob)
      atomicOr(&F.status, SSD_STATUS_BEV_OOB);
  }
}
