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
// Block-level stream compaction shared by k_label_bev and k_quad_reduce.
// A block owns a tile of SSD_TILE2 consecutive pixels. Phase A: every thread reads 16 labels (one 16 B load),
// decides which of them need per-point geometry, and the block compacts those (label, local index) pairs
// into shared memory in pixel order (warp-shuffle scan + one shared-memory round). Phase B walks the compact
// list densely: full warps, coalesced 12 B vertex gathers, no divergence on "does this point matter".
// ---------------------------------------------------------------------------------------------
#define SSD_TILE2 8192
#define SSD_ROUND2 (SSD_PT_THREADS * 16) // points per phase-A round

struct CompactList
{
  unsigned entry[SSD_TILE2]; // label << 16 | local index
  unsigned warp_tot[SSD_PT_WARPS];
  unsigned count;
};

// append the set bits of mask16 (points tid*16+b of this round) in order; all threads of the block call
__device__ __forceinline__ void compact_append(CompactList &L, unsigned mask16, const unsigned lab[4], int round, int tid)
{
  const int lane = tid & 31, warp = tid >> 5;
  const unsigned n = __popc(mask16);
  unsigned incl = n;
#pragma unroll
  for(int d = 1; d < 32; d <<= 1)
  {
    const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
    if(lane >= d)
      incl += t;
  }
  if(lane == 31)
    L.warp_tot[warp] = incl;
  __syncthreads();
  unsigned base = L.count;
#pragma unroll
  for(int w = 0; w < SSD_PT_WARPS; w++)
    base += w < warp ? L.warp_tot[w] : 0u;
  unsigned off = base + incl - n;
  unsigned m = mask16;
  while(m)
  {
    const int b = __ffs(m) - 1;
    m &= m - 1;
    const unsigned l = (lab[b >> 2] >> (8 * (b & 3))) & 0xffu;
    L.entry[off++] = (l << 16) | (unsigned)(round * SSD_ROUND2 + tid * 16 + b);
  }
  __syncthreads();
  if(tid == SSD_PT_THREADS - 1)
    L.count = base + incl; // last thread's inclusive end = new total
  __syncthreads();
}

// per-byte "label is one of the first 32 and its bit is set in amask"
__device__ __forceinline__ unsigned active_mask4(unsigned w, unsigned amask)
{
  unsigned m = 0;
#pragma unroll
  for(int j = 0; j < 4; j++)
  {
    const unsigned l = (w >> (8 * j)) & 0xffu;
    m |= (l < 32u && ((amask >> l) & 1u)) ? (1u << j) : 0u;
  }
  return m;
}

// ---------------------------------------------------------------------------------------------
// k_label_bev: the per-point segment label (PlateausExtraction::extractPlateaus, pointcloud.cpp:280-343,
// as a LUT lookup) and StairsDetector::projectToBinaryImage (:458-471) for every outlined plateau.
// Rewrites the bin codes in place as labels. grid = (ceil(N / 8192), frames).
// Phase B: x,y rows of the exact transform, BEV pixel, warp-aggregated OR into the plateau's bitmap
// (match.any on the bitmap word + redux.or: one atomic per distinct word per warp).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SSD_PT_THREADS) k_label_bev(const __grid_constant__ DevParams p, const float *__restrict__ xyz,
                                                               unsigned char *__restrict__ labels, FrameDev *__restrict__ frames,
                                                               unsigned *__restrict__ bev, size_t bm_words)
{
  __shared__ CompactList L;
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
  {
    s_oob = 0;
    L.count = 0;
  }
  const int first_outlined = F.first_outlined, K = F.n_plateaus;
  // labels first_outlined .. K-1 get a BEV image
  const unsigned amask = (K >= 32 ? 0xffffffffu : ((1u << K) - 1u)) & ~((first_outlined >= 32) ? 0xffffffffu : ((1u << first_outlined) - 1u));
  __syncthreads();

  const size_t fbase = (size_t)frame * p.N;
  const int tile0 = blockIdx.x * SSD_TILE2; // first pixel of the tile
  uint4 *lab128 = reinterpret_cast<uint4 *>(labels + fbase + tile0);
  const int tile_n = min(SSD_TILE2, p.N - tile0);

  // ---- phase A: codes -> labels, compaction of the BEV points ----
#pragma unroll
  for(int r = 0; r < SSD_TILE2 / SSD_ROUND2; r++)
  {
    const int pt0 = r * SSD_ROUND2 + tid * 16;
    unsigned lab[4] = { 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu };
    unsigned mask16 = 0;
    if(pt0 + 16 <= tile_n)
    {
      const uint4 cw = lab128[pt0 >> 4];
      const unsigned c[4] = { cw.x, cw.y, cw.z, cw.w };
#pragma unroll
      for(int i = 0; i < 4; i++)
      {
        lab[i] = (unsigned)s_lut[c[i] & 0xff] | ((unsigned)s_lut[(c[i] >> 8) & 0xff] << 8) | ((unsigned)s_lut[(c[i] >> 16) & 0xff] << 16) |
                 ((unsigned)s_lut[c[i] >> 24] << 24);
        mask16 |= active_mask4(lab[i], amask) << (4 * i);
      }
      lab128[pt0 >> 4] = make_uint4(lab[0], lab[1], lab[2], lab[3]);
    }
    else if(pt0 < tile_n)
    {
      // ragged tail of the frame (N is a multiple of 4, not necessarily of 16)
      unsigned *lab32 = reinterpret_cast<unsigned *>(labels + fbase + tile0);
      for(int i = 0; i < 4 && pt0 + 4 * i < tile_n; i++)
      {
        const unsigned c = lab32[(pt0 >> 2) + i];
        lab[i] = (unsigned)s_lut[c & 0xff] | ((unsigned)s_lut[(c >> 8) & 0xff] << 8) | ((unsigned)s_lut[(c >> 16) & 0xff] << 16) |
                 ((unsigned)s_lut[c >> 24] << 24);
        mask16 |= active_mask4(lab[i], amask) << (4 * i);
        lab32[(pt0 >> 2) + i] = lab[i];
      }
    }
    compact_append(L, mask16, lab, r, tid);
  }

  // ---- phase B: dense walk over the compacted points ----
  const int n_act = (int)L.count;
  const float *fxyz = xyz + (fbase + tile0) * 3;
  unsigned *fbev = bev + (size_t)frame * SSD_GPU_MAX_PLATEAUS * bm_words;
  const int lane = tid & 31;
  unsigned oob = 0;
  for(int base = 0; base < n_act; base += SSD_PT_THREADS)
  {
    const int i = base + tid;
    const bool live = i < n_act;
    unsigned key = 0xffffffffu, bit = 0, l = 0xffu;
    int y = 0;
    if(live)
    {
      const unsigned e = L.entry[i];
      l = e >> 16;
      const float *v = fxyz + (size_t)(e & 0xffffu) * 3;
      double wx, wy;
      camera_to_world_xy(p, __ldg(v), __ldg(v + 1), __ldg(v + 2), wx, wy);
      int x;
      if(bev_pixel(p, wx, wy, x, y))
      {
        key = l * (unsigned)bm_words + (unsigned)y * (unsigned)p.wpr + (unsigned)(x >> 5);
        bit = 1u << (x & 31);
      }
      else
        oob = 1;
    }
    // one OR per distinct bitmap word in the warp
    const unsigned grp = __match_any_sync(0xffffffffu, key);
    const unsigned bits = __reduce_or_sync(grp, bit);
    if(key != 0xffffffffu && lane == __ffs(grp) - 1)
      atomicOr(fbev + key, bits);
    // touched row range per label
    const unsigned lkey = key != 0xffffffffu ? l : 0xffu;
    const unsigned lgrp = __match_any_sync(0xffffffffu, lkey);
    const int ymin = __reduce_min_sync(lgrp, y), ymax = __reduce_max_sync(lgrp, y);
    if(lkey != 0xffu && lane == __ffs(lgrp) - 1)
    {
      if(ymin < s_rmin[lkey])
        atomicMin(&s_rmin[lkey], ymin);
      if(ymax > s_rmax[lkey])
        atomicMax(&s_rmax[lkey], ymax);
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
// z is accumulated in 2^-36 m fixed point (biased by 2^37 to stay non-negative): integer sums are order
// independent, so the result is deterministic; the error (<= 2^-37 m per point) is eight orders below the
// 0.1 mm tolerance. Per iteration the warp groups its lanes by label (match.any) and reduces each group with
// redux (single-pass segmented reduce); one lane per group adds into the warp's shared-memory accumulators.
// Point-in-quadrilateral: single-precision fast accept against the verified inner box of the step
// (quadtest_inner_box), exact QuadrilateralTest evaluation otherwise.
// Ground BEV: only the pixel columns Segmentation::detectFrontEdge can see (BottomScanner probes columns
// W/2 + 50 j; the 3x3 close reaches two columns to either side) are written.
// ---------------------------------------------------------------------------------------------
#define SSD_ZFIX_BIAS (1ll << 37)

__global__ void __launch_bounds__(SSD_PT_THREADS) k_quad_reduce(const __grid_constant__ DevParams p, const float *__restrict__ xyz,
                                                                 const unsigned char *__restrict__ labels, FrameDev *__restrict__ frames,
                                                                 unsigned *__restrict__ bev, size_t bm_words,
                                                                 unsigned long long *__restrict__ counters)
{
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ CompactList L;
  __shared__ float4 s_box[SSD_GPU_MAX_PLATEAUS]; // verified inner box of each step (cx, hx, cy, hy)
  __shared__ unsigned long long s_wsum[SSD_PT_WARPS][SSD_GPU_MAX_PLATEAUS];
  __shared__ unsigned s_wcnt[SSD_PT_WARPS][SSD_GPU_MAX_PLATEAUS];
  __shared__ int s_rmin, s_rmax;
  __shared__ unsigned s_oob, s_amask;
  QuadTestDev *s_qt = reinterpret_cast<QuadTestDev *>(s_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int frame = blockIdx.y;
  FrameDev &F = frames[frame];
  if(F.first_valid < 0)
    return; // no valid plateau: nothing is emitted (pointcloud.cpp:434)
  const int K = F.n_plateaus, ground = F.ground_index;
  {
    const bool act = tid < K && F.plat[tid].valid && F.plat[tid].quad_status == 0;
    const unsigned am = __ballot_sync(0xffffffffu, act);
    if(tid == 0)
    {
      s_amask = am;
      s_rmin = 0x7fffffff;
      s_rmax = -1;
      s_oob = 0;
      L.count = 0;
    }
    if(tid < K)
    {
      const QuadTestDev &g = F.plat[tid].qt;
      s_box[tid] = make_float4(g.ib_cx, act ? g.ib_hx : -1.f, g.ib_cy, g.ib_hy);
    }
    for(int i = tid; i < SSD_PT_WARPS * SSD_GPU_MAX_PLATEAUS; i += SSD_PT_THREADS)
    {
      (&s_wsum[0][0])[i] = 0;
      (&s_wcnt[0][0])[i] = 0;
    }
    const int words = (int)(sizeof(QuadTestDev) / 4);
    for(int i = tid; i < K * words; i += SSD_PT_THREADS)
    {
      const int k = i / words, w = i - k * words;
      reinterpret_cast<unsigned *>(&s_qt[k])[w] = reinterpret_cast<const unsigned *>(&F.plat[k].qt)[w];
    }
  }
  __syncthreads();
  const unsigned amask = s_amask;

  const size_t fbase = (size_t)frame * p.N;
  const int tile0 = blockIdx.x * SSD_TILE2;
  const uint4 *lab128 = reinterpret_cast<const uint4 *>(labels + fbase + tile0);
  const int tile_n = min(SSD_TILE2, p.N - tile0);

  // ---- phase A: compaction of the points whose label is an emitted step ----
#pragma unroll
  for(int r = 0; r < SSD_TILE2 / SSD_ROUND2; r++)
  {
    const int pt0 = r * SSD_ROUND2 + tid * 16;
    unsigned lab[4] = { 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu };
    unsigned mask16 = 0;
    if(pt0 + 16 <= tile_n)
    {
      const uint4 lw = __ldg(lab128 + (pt0 >> 4));
      lab[0] = lw.x;
      lab[1] = lw.y;
      lab[2] = lw.z;
      lab[3] = lw.w;
    }
    else if(pt0 < tile_n)
    {
      const unsigned *lab32 = reinterpret_cast<const unsigned *>(labels + fbase + tile0);
      for(int i = 0; i < 4 && pt0 + 4 * i < tile_n; i++)
        lab[i] = __ldg(lab32 + (pt0 >> 2) + i);
    }
#pragma unroll
    for(int i = 0; i < 4; i++)
      if(__vcmpltu4(lab[i], 0x20202020u)) // any plateau label in this word at all?
        mask16 |= active_mask4(lab[i], amask) << (4 * i);
    compact_append(L, mask16, lab, r, tid);
  }

  // ---- phase B ----
  const int n_act = (int)L.count;
  const float *fxyz = xyz + (fbase + tile0) * 3;
  unsigned *gbev = ground >= 0 ? bev + ((size_t)frame * SSD_GPU_MAX_PLATEAUS + ground) * bm_words : nullptr;
  unsigned oob = 0, n_fast = 0, n_slow = 0;
  int rmin = 0x7fffffff, rmax = -1;
  // ground BEV column filter: pixel column u needed iff (u - (W/2 - 2)) mod 50 in [0, 5)
  const float sxf = (float)p.x_to_image, kxf = (float)(-p.x_min * p.x_to_image) - (float)(p.W / 2 - 2);
  const float dux = (float)(p.x_to_image * 1.0001) , du0 = (float)(8.0 * p.W / 16777216.0) + 1e-3f;

  for(int base = 0; base < n_act; base += SSD_PT_THREADS)
  {
    const int i = base + tid;
    unsigned l = 0xffu;
    unsigned long long zf = 0;
    bool inside = false;
    if(i < n_act)
    {
      const unsigned e = L.entry[i];
      const unsigned le = e >> 16;
      const float *v = fxyz + (size_t)(e & 0xffffu) * 3;
      const float fx = __ldg(v), fy = __ldg(v + 1), fz = __ldg(v + 2);
      // fast accept: single-precision position inside the verified inner box by more than its error bound
      // (same bound as point_code_filtered: |w^ - w_ref| <= eps = E1 * max|p| + E0)
      const float4 bx = s_box[le];
      const float eps = fmaf(p.E1, fmaxf(fmaxf(fabsf(fx), fabsf(fy)), fabsf(fz)), p.E0);
      const float wxf = fmaf(p.af[2], fz, fmaf(p.af[1], fy, fmaf(p.af[0], fx, p.bf[0])));
      const float wyf = fmaf(p.af[5], fz, fmaf(p.af[4], fy, fmaf(p.af[3], fx, p.bf[1])));
      const bool fast = fmaxf(fabsf(wxf - bx.x) - bx.y, fabsf(wyf - bx.z) - bx.w) < -eps;
      double wx = 0, wy = 0;
      bool have_xy = false;
      inside = fast;
      n_fast += fast;
      n_slow += !fast;
      if(!fast)
      {
        camera_to_world_xy(p, fx, fy, fz, wx, wy);
        have_xy = true;
        inside = quadtest_within(s_qt[le], wx, wy);
      }
      if(inside)
      {
        l = le;
        zf = (unsigned long long)(z_to_fix(camera_to_world_z(p, fx, fy, fz)) + SSD_ZFIX_BIAS);
        if((int)le == ground)
        {
          // is the pixel column one the front-edge scanner can see? (f32 estimate, conservative margin)
          const float uf = fmaf(wxf, sxf, kxf);              // pixel x coordinate relative to W/2 - 2
          const float q50 = floorf(uf * 0.02f);
          const float tcol = fmaf(q50, -50.f, uf);            // uf mod 50 (approximately, in [-tiny, 50+tiny])
          const float dcol = fmaf(eps, dux, du0);
          // (a pixel that rounds to column W wraps to column 0 of the next row: keep the right edge too)
          if(tcol < 5.f + dcol || tcol > 50.f - dcol || uf > (float)(p.W - 2 - (p.W / 2 - 2)))
          {
            if(!have_xy)
              camera_to_world_xy(p, fx, fy, fz, wx, wy);
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
    }
    // single-pass segmented reduce over the warp: lanes grouped by label
    const unsigned grp = __match_any_sync(0xffffffffu, l);
    const unsigned lo = __reduce_add_sync(grp, (unsigned)(zf & 0xfffffu));          // 20 + 19 bits: sums of 32 fit in 32 bits
    const unsigned hi = __reduce_add_sync(grp, (unsigned)(zf >> 20));
    if(l != 0xffu && lane == __ffs(grp) - 1)
    {
      s_wsum[warp][l] += ((unsigned long long)hi << 20) + lo; // one lane per label per warp: no race
      s_wcnt[warp][l] += __popc(grp);
    }
    __syncwarp();
  }
  {
    n_fast = __reduce_add_sync(0xffffffffu, n_fast);
    n_slow = __reduce_add_sync(0xffffffffu, n_slow);
    if(lane == 0 && counters && (n_fast | n_slow))
    {
      atomicAdd(counters + 1, (unsigned long long)n_fast);
      atomicAdd(counters + 2, (unsigned long long)n_slow);
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
  if(tid < K)
  {
    unsigned long long s = 0;
    unsigned n = 0;
#pragma unroll
    for(int w = 0; w < SSD_PT_WARPS; w++)
    {
      s += s_wsum[w][tid];
      n += s_wcnt[w][tid];
    }
    if(n)
    {
      atomicAdd(&F.plat[tid].sum_fix, s);
      atomicAdd(&F.plat[tid].n_in_quad, n);
    }
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
