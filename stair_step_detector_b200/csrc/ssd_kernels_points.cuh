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
  unsigned run_code = 0xffffffffu, run_n = 0, exact = 0;

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
      if(unc[0] | unc[1] | unc[2] | unc[3])
      {
#pragma unroll
        for(int j = 0; j < 4; j++)
          if(unc[j])
          {
            c[j] = point_code_slow(p, v.x[j], v.y[j], v.z[j]);
            exact++;
          }
      }
      codes32[q] = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
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

// warp-aggregated row-range tracking + OR into a BEV bitmap word
__device__ __forceinline__ void bev_set(const DevParams &p, unsigned *__restrict__ bm, double wx, double wy, int *s_rmin, int *s_rmax, int k,
                                        unsigned &oob)
{
  int ix, iy;
  world_to_image(p, wx, wy, ix, iy);
  const long long off = (long long)iy * p.W + ix; // cv::Mat::ptr(y, x) arithmetic, no bounds check (pointcloud.cpp:468)
  if(off < 0 || off >= (long long)p.N)
  {
    oob = 1;
    return;
  }
  const int y = (int)(off / p.W), x = (int)(off - (long long)y * p.W);
  atomicOr(bm + (size_t)y * p.wpr + (x >> 5), 1u << (x & 31));
  if(y < s_rmin[k])
    atomicMin(&s_rmin[k], y);
  if(y > s_rmax[k])
    atomicMax(&s_rmax[k], y);
}

// ---------------------------------------------------------------------------------------------
// k_label_bev: the per-point segment label (PlateausExtraction::extractPlateaus, pointcloud.cpp:280-343,
// as a LUT lookup) and StairsDetector::projectToBinaryImage (:458-471) for every outlined plateau.
// Rewrites the bin codes in place as labels. grid = (tiles_per_frame, frames).
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

#pragma unroll
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
          double wx, wy, wz;
          camera_to_world(p, v.x[j], v.y[j], v.z[j], wx, wy, wz);
          bev_set(p, fbev + (size_t)l[j] * bm_words, wx, wy, s_rmin, s_rmax, (int)l[j], oob);
        }
    }
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

// ---------------------------------------------------------------------------------------------
// k_quad_reduce: StairsDetector::getPointsInQuadrilateral + calcAverageZ (pointcloud.cpp:560-581) for the
// ground and every valid plateau, and the ground's BEV image (calcGround, :530-531).
// z is accumulated in 2^-36 m fixed point: integer sums are order independent, so the result is
// deterministic; the error (<= 2^-37 m per point) is eight orders below the 0.1 mm tolerance.
// ---------------------------------------------------------------------------------------------
template<int ITERS>
__global__ void __launch_bounds__(SSD_PT_THREADS) k_quad_reduce(const __grid_constant__ DevParams p, const float *__restrict__ xyz,
                                                                 const unsigned char *__restrict__ labels, FrameDev *__restrict__ frames,
                                                                 unsigned *__restrict__ bev, size_t bm_words)
{
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ unsigned long long s_sum[SSD_GPU_MAX_PLATEAUS];
  __shared__ unsigned s_cnt[SSD_GPU_MAX_PLATEAUS];
  __shared__ int s_active[SSD_GPU_MAX_PLATEAUS];
  __shared__ int s_rmin, s_rmax;
  __shared__ unsigned s_oob;
  QuadTestDev *s_qt = reinterpret_cast<QuadTestDev *>(s_raw);
  const int tid = threadIdx.x;
  const int frame = blockIdx.y;
  FrameDev &F = frames[frame];
  if(F.first_valid < 0)
    return; // no valid plateau: nothing is emitted (pointcloud.cpp:434)
  const int K = F.n_plateaus, ground = F.ground_index;
  if(tid < SSD_GPU_MAX_PLATEAUS)
  {
    s_sum[tid] = 0;
    s_cnt[tid] = 0;
    s_active[tid] = tid < K && F.plat[tid].valid && F.plat[tid].quad_status == 0;
  }
  if(tid == 0)
  {
    s_rmin = 0x7fffffff;
    s_rmax = -1;
    s_oob = 0;
  }
  {
    const unsigned *src = reinterpret_cast<const unsigned *>(&F.plat[0]);
    (void)src;
    for(int k = 0; k < K; k++)
    {
      const unsigned *g = reinterpret_cast<const unsigned *>(&F.plat[k].qt);
      unsigned *d = reinterpret_cast<unsigned *>(&s_qt[k]);
      for(int i = tid; i < (int)(sizeof(QuadTestDev) / 4); i += SSD_PT_THREADS)
        d[i] = g[i];
    }
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
#pragma unroll
  for(int it = 0; it < ITERS; it++)
  {
    const int q = (blockIdx.x * ITERS + it) * SSD_PT_THREADS + tid;
    if(q >= nquads)
      continue;
    const unsigned lw = __ldg(lab32 + q);
    bool any = false;
#pragma unroll
    for(int j = 0; j < 4; j++)
    {
      const unsigned l = (lw >> (8 * j)) & 0xff;
      any |= l < SSD_GPU_MAX_PLATEAUS && s_active[l];
    }
    if(!any)
      continue;
    const Quad4 v = load_quad(xyz4, q);
#pragma unroll
    for(int j = 0; j < 4; j++)
    {
      const unsigned l = (lw >> (8 * j)) & 0xff;
      if(!(l < SSD_GPU_MAX_PLATEAUS && s_active[l]))
        continue;
      double wx, wy, wz;
      camera_to_world(p, v.x[j], v.y[j], v.z[j], wx, wy, wz);
      if(!quadtest_within(s_qt[l], wx, wy))
        continue;
      if((int)l != acc_k)
      {
        if(acc_n)
        {
          atomicAdd(&s_sum[acc_k], (unsigned long long)acc);
          atomicAdd(&s_cnt[acc_k], acc_n);
        }
        acc = 0;
        acc_n = 0;
        acc_k = (int)l;
      }
      acc += z_to_fix(wz);
      acc_n++;
      if((int)l == ground)
      {
        int ix, iy;
        world_to_image(p, wx, wy, ix, iy);
        const long long off = (long long)iy * p.W + ix;
        if(off < 0 || off >= (long long)p.N)
          oob = 1;
        else
        {
          const int y = (int)(off / p.W), x = (int)(off - (long long)y * p.W);
          atomicOr(gbev + (size_t)y * p.wpr + (x >> 5), 1u << (x & 31));
          rmin = min(rmin, y);
          rmax = max(rmax, y);
        }
      }
    }
  }
  if(acc_n)
  {
    atomicAdd(&s_sum[acc_k], (unsigned long long)acc);
    atomicAdd(&s_cnt[acc_k], acc_n);
  }
  if(rmax >= 0)
  {
    atomicMin(&s_rmin, rmin);
    atomicMax(&s_rmax, rmax);
  }
  if(oob)
    s_oob = 1;
  __syncthreads();
  if(tid < K && s_cnt[tid])
  {
    atomicAdd(&F.plat[tid].sum_fix, s_sum[tid]);
    atomicAdd(&F.plat[tid].n_in_quad, s_cnt[tid]);
  }
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
