// ssd_kernels_records.cuh -- the record chain (EXPERIMENTAL, SSD_GPU_PATH=records): the three point passes of the classic chain
// (ssd_kernels_points.cuh) with the vertices read ONCE. Parity-tested like the default chain; measured SLOWER (166 k frames/s
// against 268 k: more instructions per point than the classic passes, DESIGN.md section 6 "Round 2").
//   k_transform_rec : k_transform_bin (z>0, CameraToWorld, range filter, height bin, histogram; pointcloud.cpp:122-178,
//                     194-204) that also leaves a 4-byte record per in-range point {BEV pixel (pointcloud.cpp:79-83), parity of
//                     the bin, height offset inside the bin} -- the phase-1 body of the resident-frame chain (fs_phase1)
//   k_peaks         : unchanged
//   k_label_rec     : per-point segment labels and BEV bitmaps of the outlined plateaus (pointcloud.cpp:280-343, 458-471) from
//                     the 1-byte codes and the records alone: no vertex, no transform
//   k_quad_rec      : getPointsInQuadrilateral + calcAverageZ (pointcloud.cpp:560-581) from labels and records: a point whose
//                     BEV pixel lies in a pixel box inside the verified inner box of its step's QuadrilateralTest is inside
//                     (two integer comparisons); only the fringe along the quadrilateral's edges is re-read as vertices
//   (k_label_sum / k_quad_sum: the summary-based variants shared with the resident-frame chain)
// DRAM traffic per point (ncu, profiles/r02_records_ncu_full.csv): 15.0 | 2.6 | 3.3 = 21 B against the classic chain's 24.
#pragma once
#include "ssd_kernels_stream.cuh"

struct RecAcc // block-level accumulators of k_transform_rec (the fields fs_phase1 touches)
{
  unsigned hist[SSD_BINS_PAD];
  unsigned n_exact, n_def;
};

template<int ITERS>
__global__ void __launch_bounds__(SSD_PT_THREADS) k_transform_rec(const __grid_constant__ DevParams p, const float *__restrict__ xyz,
                                                                   unsigned char *__restrict__ codes, uint4 *__restrict__ recs,
                                                                   FrameDev *__restrict__ frames)
{
  extern __shared__ __align__(128) unsigned char s_dyn[]; // ITERS stages of SSD_TB_STAGE_BYTES
  __shared__ __align__(8) unsigned long long s_bar[ITERS];
  __shared__ RecAcc A;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int frame = blockIdx.y;
  const size_t fbase = (size_t)frame * p.N;
  const int nquads = p.N >> 2;
  const int qb = blockIdx.x * (ITERS * SSD_PT_THREADS); // first 4-point word of the block
  const unsigned stage_sa = (unsigned)__cvta_generic_to_shared(s_dyn);
  const unsigned bar_sa = (unsigned)__cvta_generic_to_shared(s_bar);
  if(tid == 0)
  {
#pragma unroll
    for(int it = 0; it < ITERS; it++)
      mbar_init(bar_sa + it * 8, 1);
    mbar_fence_init();
    const unsigned char *gsrc = reinterpret_cast<const unsigned char *>(xyz + fbase * 3) + (size_t)qb * 48;
#pragma unroll
    for(int it = 0; it < ITERS; it++)
    {
      const int nq = min(SSD_PT_THREADS, nquads - (qb + it * SSD_PT_THREADS));
      if(nq > 0)
      {
        mbar_expect_tx(bar_sa + it * 8, (unsigned)nq * 48u);
        bulk_g2s(stage_sa + it * SSD_TB_STAGE_BYTES, gsrc + (size_t)it * SSD_TB_STAGE_BYTES, (unsigned)nq * 48u, bar_sa + it * 8);
      }
    }
    A.n_exact = 0;
    A.n_def = 0;
  }
  A.hist[tid] = 0; // SSD_PT_THREADS == SSD_BINS_PAD
  __syncthreads();

  const SrcVertices src = { xyz };
  unsigned *code32 = reinterpret_cast<unsigned *>(codes + fbase);
  uint4 *rec16 = recs + (size_t)frame * nquads;
#pragma unroll
  for(int it = 0; it < ITERS; it++)
  {
    const int q = qb + it * SSD_PT_THREADS + tid;
    if(qb + it * SSD_PT_THREADS < nquads) // block-uniform: the stage was requested (N % 128 == 0: whole warps)
    {
      mbar_wait(bar_sa + it * 8, 0);
      fs_phase1(p, src, s_dyn + it * SSD_TB_STAGE_BYTES + warp * (SSD_FS_STEP_PX * 12), code32 + q, rec16 + q, A, (unsigned)(q >> 5), lane);
    }
  }
  __syncthreads();
  const unsigned sum = A.hist[tid];
  if(sum)
    atomicAdd(&frames[frame].hist[tid], sum);
  if(tid == 0)
  {
    if(A.n_exact)
      atomicAdd(&frames[frame].n_exact_bin, A.n_exact);
    if(A.n_def)
      atomicAdd(&frames[frame].n_def_bev, A.n_def);
  }
}

// the same pass on z16 depth frames (2 bytes per point in; the deprojected vertices live in registers only)
template<int ITERS>
__global__ void __launch_bounds__(SSD_PT_THREADS) k_transform_rec_depth(const __grid_constant__ DevParams p, const SrcDepth src,
                                                                         unsigned char *__restrict__ codes, uint4 *__restrict__ recs,
                                                                         FrameDev *__restrict__ frames)
{
  __shared__ RecAcc A;
  const int tid = threadIdx.x, lane = tid & 31;
  const int frame = blockIdx.y;
  const size_t fbase = (size_t)frame * p.N;
  const int nquads = p.N >> 2;
  const int qb = blockIdx.x * (ITERS * SSD_PT_THREADS);
  A.hist[tid] = 0;
  if(tid == 0)
  {
    A.n_exact = 0;
    A.n_def = 0;
  }
  __syncthreads();
  unsigned *code32 = reinterpret_cast<unsigned *>(codes + fbase);
  uint4 *rec16 = recs + (size_t)frame * nquads;
#pragma unroll 1
  for(int it = 0; it < ITERS; it++)
  {
    const int q = qb + it * SSD_PT_THREADS + tid;
    if(qb + it * SSD_PT_THREADS < nquads)
    {
      const unsigned step = (unsigned)(q >> 5);
      fs_phase1(p, src, reinterpret_cast<const unsigned char *>(src.z16 + fbase + (size_t)step * SSD_FS_STEP_PX), code32 + q, rec16 + q, A, step, lane);
    }
  }
  __syncthreads();
  const unsigned sum = A.hist[tid];
  if(sum)
    atomicAdd(&frames[frame].hist[tid], sum);
  if(tid == 0)
  {
    if(A.n_exact)
      atomicAdd(&frames[frame].n_exact_bin, A.n_exact);
    if(A.n_def)
      atomicAdd(&frames[frame].n_def_bev, A.n_def);
  }
}

struct LabelSumShared // block-level state of k_label_sum (the fields fs_phase2 touches)
{
  unsigned short lut[SSD_BINS_PAD];
  int rmin[SSD_GPU_MAX_PLATEAUS], rmax[SSD_GPU_MAX_PLATEAUS];
  unsigned oob, n_def;
};

// grid = (blocks per frame, frames); every warp walks its own 128-point sub-steps of the frame (interleaved over the warps of
// the frame's blocks); the next sub-step's code word is requested while the current one is worked on.
template<class SRC>
__global__ void __launch_bounds__(SSD_PT_THREADS) k_label_sum(const __grid_constant__ DevParams p, const SRC src, unsigned char *__restrict__ labels,
                                                               const uint4 *__restrict__ recs, FrameDev *__restrict__ frames,
                                                               unsigned *__restrict__ bev, size_t bm_words, GroupSum *__restrict__ sums)
{
  __shared__ LabelSumShared S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int frame = blockIdx.y;
  FrameDev &F = frames[frame];
  const int nquads = p.N >> 2, nsteps = p.gs_steps;
  const unsigned *code32 = reinterpret_cast<const unsigned *>(labels + (size_t)frame * p.N);
  const uint4 *rec16 = recs + (size_t)frame * nquads;
  S.lut[tid] = F.lut16[tid]; // SSD_PT_THREADS == SSD_BINS_PAD
  if(tid < SSD_GPU_MAX_PLATEAUS)
  {
    S.rmin[tid] = 0x7fffffff;
    S.rmax[tid] = -1;
  }
  if(tid == 0)
  {
    S.oob = 0;
    S.n_def = 0;
  }
  __syncthreads();
  FsParams a;
  a.n_frames = 0;
  a.d_raw = a.d_rec = a.flags = 0;
  a.recs = nullptr;
  a.done = nullptr;
  a.sums = sums;
  a.prof = nullptr;
  a.prof_warp_off = 0;
  const int stride = gridDim.x * SSD_PT_WARPS;
  int st = blockIdx.x * SSD_PT_WARPS + warp;
  unsigned cw = st < nsteps ? __ldcs(code32 + (size_t)st * 32 + lane) : 0xffffffffu;
  for(; st < nsteps; st += stride)
  {
    const int nx = st + stride;
    const unsigned cwn = nx < nsteps ? __ldcs(code32 + (size_t)nx * 32 + lane) : 0xffffffffu;
    uint4 rv = make_uint4(0u, 0u, 0u, 0u);
    if((cw & 0xfefefefeu) != 0xfefefefeu)
      rv = __ldcs(rec16 + (size_t)st * 32 + lane);
    fs_phase2(p, src, a, cw, rv, S.lut, S, (unsigned)frame, (unsigned)st, labels, bev, (unsigned)bm_words, lane);
    cw = cwn;
  }
  __syncthreads();
  if(tid < SSD_GPU_MAX_PLATEAUS && S.rmax[tid] >= 0)
  {
    atomicMin(&F.plat[tid].row_min, S.rmin[tid]);
    atomicMax(&F.plat[tid].row_max, S.rmax[tid]);
  }
  if(tid == 0 && S.oob)
    atomicOr(&F.status, SSD_STATUS_BEV_OOB);
}

// ---------------------------------------------------------------------------------------------
// k_label_rec: labels + BEV bitmaps of the outlined plateaus from codes and records (no summaries: k_quad_rec reads the
// records itself). Per 128-point sub-step: code word -> LUT -> label word (stored only where it differs from the code: codes
// 254 / 255 are their own labels); the lanes holding points of outlined plateaus load their four records and set the BEV bits.
// ---------------------------------------------------------------------------------------------
template<class SRC>
__global__ void __launch_bounds__(SSD_PT_THREADS) k_label_rec(const __grid_constant__ DevParams p, const SRC src, unsigned char *__restrict__ labels,
                                                               const uint4 *__restrict__ recs, FrameDev *__restrict__ frames,
                                                               unsigned *__restrict__ bev, size_t bm_words)
{
  __shared__ LabelSumShared S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int frame = blockIdx.y;
  FrameDev &F = frames[frame];
  const int nquads = p.N >> 2, nsteps = p.gs_steps;
  unsigned *lab32 = reinterpret_cast<unsigned *>(labels + (size_t)frame * p.N);
  const uint4 *rec16 = recs + (size_t)frame * nquads;
  S.lut[tid] = F.lut16[tid]; // SSD_PT_THREADS == SSD_BINS_PAD
  if(tid < SSD_GPU_MAX_PLATEAUS)
  {
    S.rmin[tid] = 0x7fffffff;
    S.rmax[tid] = -1;
  }
  if(tid == 0)
  {
    S.oob = 0;
    S.n_def = 0;
  }
  __syncthreads();
  const int bx = p.rec_bx;
  const unsigned mx = (1u << bx) - 1u, my = (1u << p.rec_by) - 1u;
  const unsigned bmw = (unsigned)bm_words, wpr = (unsigned)p.wpr;
  unsigned *fbev = bev + (size_t)frame * SSD_GPU_MAX_PLATEAUS * bm_words;
  unsigned rl = 0xffu; // label of the lane's current row-range run
  int rlo = 0x7fffffff, rhi = -1;
  const int stride = gridDim.x * SSD_PT_WARPS;
  int st = blockIdx.x * SSD_PT_WARPS + warp;
  unsigned cw = st < nsteps ? __ldcs(lab32 + (size_t)st * 32 + lane) : 0xffffffffu;
  for(; st < nsteps; st += stride)
  {
    const int nx = st + stride;
    const unsigned cwn = nx < nsteps ? __ldcs(lab32 + (size_t)nx * 32 + lane) : 0xffffffffu;
    if((cw & 0xfefefefeu) != 0xfefefefeu)
    {
      const unsigned e0 = S.lut[cw & 0xffu], e1 = S.lut[(cw >> 8) & 0xffu], e2 = S.lut[(cw >> 16) & 0xffu], e3 = S.lut[cw >> 24];
      const unsigned lab = (e0 & 0xffu) | ((e1 & 0xffu) << 8) | ((e2 & 0xffu) << 16) | (e3 << 24);
      const unsigned ol = ((e0 >> 8) & 1u) | ((e1 >> 7) & 2u) | ((e2 >> 6) & 4u) | ((e3 >> 5) & 8u);
      lab32[(size_t)st * 32 + lane] = lab;
      if(ol)
      {
        const uint4 rv = __ldcs(rec16 + (size_t)st * 32 + lane);
        const unsigned r[4] = { rv.x, rv.y, rv.z, rv.w };
#pragma unroll
        for(int j = 0; j < 4; j++)
          if((ol >> j) & 1u)
          {
            const unsigned l = (lab >> (8 * j)) & 0xffu;
            int ix = (int)(r[j] & mx), iy = (int)((r[j] >> bx) & my);
            if(iy >= p.H)
            {
              // pixel outside the image (the x == W wrap / past the end, pointcloud.cpp:81,468): the reference's unchecked
              // write, from the vertex itself
              const typename SrcTraits<SRC>::Frame FR = src_frame(src, p, (size_t)frame * p.N);
              float fx, fy, fz;
              point_load(FR, (unsigned)st * SSD_FS_STEP_PX + (unsigned)lane * 4u + (unsigned)j, fx, fy, fz);
              double wx, wy;
              camera_to_world_xy(p, fx, fy, fz, wx, wy);
              if(!bev_pixel(p, wx, wy, ix, iy))
              {
                S.oob = 1;
                continue;
              }
            }
            atomicOr(fbev + (l * bmw + (unsigned)iy * wpr + ((unsigned)ix >> 5)), 1u << (ix & 31));
            if(l != rl)
            {
              if(rhi >= 0)
              {
                atomicMin(&S.rmin[rl], rlo);
                atomicMax(&S.rmax[rl], rhi);
              }
              rl = l;
              rlo = 0x7fffffff;
              rhi = -1;
            }
            rlo = min(rlo, iy);
            rhi = max(rhi, iy);
          }
      }
    }
    cw = cwn;
  }
  if(rhi >= 0)
  {
    atomicMin(&S.rmin[rl], rlo);
    atomicMax(&S.rmax[rl], rhi);
  }
  __syncthreads();
  if(tid < SSD_GPU_MAX_PLATEAUS && S.rmax[tid] >= 0)
  {
    atomicMin(&F.plat[tid].row_min, S.rmin[tid]);
    atomicMax(&F.plat[tid].row_max, S.rmax[tid]);
  }
  if(tid == 0 && S.oob)
    atomicOr(&F.status, SSD_STATUS_BEV_OOB);
}

// ---------------------------------------------------------------------------------------------
// k_quad_rec: getPointsInQuadrilateral + calcAverageZ (pointcloud.cpp:560-581) and the ground's BEV image (:530-531) from
// labels and records. The verified inner box of each step's QuadrilateralTest (quadtest_inner_box: isPointWithin() is true on
// the whole box) is turned into a box of BEV PIXELS whose world rectangles lie inside it: a point whose pixel is in that box
// is inside the quadrilateral -- two integer comparisons on its record, no vertex, no transform; its height enters the sum as
// (bin code, offset) integers. A pixel beyond the reject box is outside. Only the points in between -- the fringe along the
// quadrilateral's edges -- are re-read as vertices and decided by the filtered / exact test of k_quad_reduce (qr_dense).
// ---------------------------------------------------------------------------------------------
struct PixBox
{
  int ix0, ixw, iy0, iyh;     // inside <=> (unsigned)(ix - ix0) <= ixw && (unsigned)(iy - iy0) <= iyh   (none: ix0 = 2^30, ixw = 0)
  int rx0, rxw, ry0, ryh;     // possibly inside the bounding box <=> (unsigned)(ix - rx0) <= rxw && (unsigned)(iy - ry0) <= ryh
  int hmin, pad0, pad1, pad2; // lower bin of the plateau's band: a point's bin code is hmin + ((parity bit of its record ^ hmin) & 1)
};
struct QuadRecShared
{
  PixBox pb[SSD_GPU_MAX_PLATEAUS];
  int sd[SSD_GPU_MAX_PLATEAUS];
  unsigned sc[SSD_GPU_MAX_PLATEAUS], sn[SSD_GPU_MAX_PLATEAUS];
};

template<class SRC>
__global__ void __launch_bounds__(SSD_PT_THREADS, SSD_QR_MINB) k_quad_rec(const __grid_constant__ DevParams p, const SRC src,
                                                                          const unsigned char *__restrict__ labels, FrameDev *__restrict__ frames,
                                                                          unsigned *__restrict__ bev, size_t bm_words, const uint4 *__restrict__ recs)
{
  __shared__ QuadReduceShared S;
  __shared__ QuadRecShared Q;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int frame = blockIdx.y;
  FrameDev &F = frames[frame];
  const unsigned amask = F.quad_amask;
  if(amask == 0u)
    return; // no valid plateau: nothing is emitted (pointcloud.cpp:434)
  const int ground = F.ground_index;
  const size_t fbase = (size_t)frame * p.N;
  const unsigned *lab32 = reinterpret_cast<const unsigned *>(labels + fbase);
  const uint4 *rec16 = recs + (size_t)frame * (p.N >> 2);
  unsigned *gbev = ground >= 0 ? bev + ((size_t)frame * SSD_GPU_MAX_PLATEAUS + ground) * bm_words : nullptr;
  const int nquads = p.N >> 2;
  int wt, wt_end, wt_stride;
  warp_tile_range((p.N + SSD_WT_PX - 1) / SSD_WT_PX, warp, wt, wt_end, wt_stride);

  qr_init(S, F, amask, tid);
  if(tid < SSD_GPU_MAX_PLATEAUS)
  {
    Q.sd[tid] = 0;
    Q.sc[tid] = 0;
    Q.sn[tid] = 0;
    PixBox b;
    b.ix0 = b.iy0 = 1 << 30;
    b.ixw = b.iyh = 0;
    b.rx0 = b.ry0 = 1 << 30;
    b.rxw = b.ryh = 0; // not live: every point "beyond the reject box"
    b.hmin = F.plat[tid].hmin;
    b.pad0 = b.pad1 = b.pad2 = 0;
    if((amask >> tid) & 1u)
    {
      const float4 ib = F.qf[tid].ib; // cx, hx, cy, hy of the verified box (f32, rounded inwards); hx < 0: none
      const float4 rj = F.qf[tid].rj; // reject half widths about the centre of ibe
      const float4 ibe = F.qf[tid].ibe;
      const double m = 1e-7;          // metres: far above the rounding of the pixel computation (1e-13), far below a pixel
      if(ib.y > 0.f && ib.w > 0.f)
      {
        // pixel ix covers x in [x_min + ix / sx, x_min + (ix + 1) / sx], pixel iy covers y in [y_max - (iy + 1) / sy, y_max - iy / sy]
        const double xlo = (double)ib.x - (double)ib.y + m, xhi = (double)ib.x + (double)ib.y - m;
        const double ylo = (double)ib.z - (double)ib.w + m, yhi = (double)ib.z + (double)ib.w - m;
        const int ix0 = (int)ceil((xlo - p.x_min) * p.x_to_image + 1e-6), ix1 = (int)floor((xhi - p.x_min) * p.x_to_image - 1e-6) - 1;
        const int iy0 = (int)ceil((p.y_max - yhi) * p.y_to_image + 1e-6), iy1 = (int)floor((p.y_max - ylo) * p.y_to_image - 1e-6) - 1;
        if(ix1 >= ix0 && iy1 >= iy0)
        {
          b.ix0 = ix0;
          b.ixw = ix1 - ix0;
          b.iy0 = iy0;
          b.iyh = iy1 - iy0;
        }
      }
      // candidates: pixels that may touch the bounding box of the quadrilateral (everything else is certainly outside)
      const double rxl = (double)ibe.x - (double)rj.x, rxh = (double)ibe.x + (double)rj.x;
      const double ryl = (double)ibe.y - (double)rj.y, ryh = (double)ibe.y + (double)rj.y;
      const double c0 = floor((rxl - p.x_min) * p.x_to_image) - 1.0, c1 = ceil((rxh - p.x_min) * p.x_to_image) + 1.0;
      const double r0 = floor((p.y_max - ryh) * p.y_to_image) - 1.0, r1 = ceil((p.y_max - ryl) * p.y_to_image) + 1.0;
      b.rx0 = (int)fmax(c0, -1.0);
      b.rxw = (int)fmin(c1, (double)p.W + 1.0) - b.rx0;
      b.ry0 = (int)fmax(r0, -1.0);
      b.ryh = (int)fmin(r1, (double)p.H + 1.0) - b.ry0;
    }
    Q.pb[tid] = b;
  }
  __syncthreads();

  unsigned short *act = S.L.act[warp];
  unsigned *labs = S.L.lab[warp];
  QrWarp W = { 0xffu, 0u, 0u, 0u, 0ull, 0x7fffffff, -1 };
  const typename SrcTraits<SRC>::Frame FR = src_frame(src, p, fbase);
  unsigned run_l = 0xffu, run_n = 0, run_c = 0;
  int run_d = 0;
  const int bx = p.rec_bx, zsh = 32 - p.rec_zbits, pbit = p.rec_bx + p.rec_by;
  const unsigned mx = (1u << bx) - 1u, my = (1u << p.rec_by) - 1u;
  unsigned cl = 0xffu; // label whose pixel boxes sit in b
  PixBox b = Q.pb[0];

  for(; wt < wt_end; wt += wt_stride)
  {
    const unsigned wbase = (unsigned)wt * (SSD_WT_PX / 4);
    unsigned n = 0;
    unsigned labw[SSD_WT_WORDS];
#pragma unroll
    for(int it = 0; it < SSD_WT_WORDS; it++)
    {
      const int q = wt * (SSD_WT_PX / 4) + it * 32 + lane;
      labw[it] = q < nquads ? __ldg(lab32 + q) : 0xffffffffu;
    }
    if(wt + wt_stride < wt_end)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(lab32 + (size_t)(wt + wt_stride) * (SSD_WT_PX / 4) + lane * 8));
#pragma unroll
    for(int it = 0; it < SSD_WT_WORDS; it++)
    {
      const unsigned lw = labw[it];
      const unsigned am4 = (((~lw >> 7) & 0x01010101u) * 0x10204080u) >> 28; // bit j <-> byte j < 128 (a plateau label)
      unsigned und = 0; // points that need the vertex
      if(am4)
      {
        const unsigned widx = (unsigned)(it * 32 + lane);
        // label of the word's first plateau point; the word is "uniform" when all its plateau points carry it (image rows are
        // iso-height: almost every word is). Mixed words go to the dense pass whole.
        const unsigned l0 = (lw >> (8 * (__ffs(am4) - 1) & 31)) & 0x1fu;
        const unsigned bytes = ((am4 * 0x00204081u) & 0x01010101u) * 0xffu;
        if(((lw ^ (l0 * 0x01010101u)) & bytes) != 0u)
          und = am4;
        else if((amask >> l0) & 1u)
        {
          if(l0 != cl)
          {
            b = Q.pb[l0];
            cl = l0;
          }
          const uint4 rv = __ldg(rec16 + wbase + widx);
          const unsigned r[4] = { rv.x, rv.y, rv.z, rv.w };
          int dsum = 0, ixs[4], iys[4];
          unsigned csum = 0, insm = 0;
#pragma unroll
          for(int j = 0; j < 4; j++)
          {
            const int ix = (int)(r[j] & mx), iy = (int)((r[j] >> bx) & my);
            ixs[j] = ix;
            iys[j] = iy;
            const bool a_j = (am4 >> j) & 1u;
            const bool ins = a_j && (unsigned)(ix - b.ix0) <= (unsigned)b.ixw && (unsigned)(iy - b.iy0) <= (unsigned)b.iyh;
            // not inside the pixel box: a candidate unless its pixel lies beyond the reject box (a record without a pixel,
            // iy == H, is always a candidate)
            const bool cand = a_j && !ins && (iy >= p.H || ((unsigned)(ix - b.rx0) <= (unsigned)b.rxw && (unsigned)(iy - b.ry0) <= (unsigned)b.ryh));
            dsum += ins ? ((int)r[j] >> zsh) : 0;
            csum += ins ? ((unsigned)b.hmin + (((r[j] >> pbit) ^ (unsigned)b.hmin) & 1u)) : 0u;
            insm |= ins ? (1u << j) : 0u;
            und |= cand ? (1u << j) : 0u;
          }
          if(insm)
          {
            if(l0 != run_l)
            {
              if(run_n)
              {
                atomicAdd(&Q.sd[run_l], run_d);
                atomicAdd(&Q.sc[run_l], run_c);
                atomicAdd(&Q.sn[run_l], run_n);
              }
              run_l = l0;
              run_d = 0;
              run_c = 0;
              run_n = 0;
            }
            run_d += dsum;
            run_c += csum;
            run_n += (unsigned)__popc(insm);
            if((int)l0 == ground)
            {
#pragma unroll
              for(int j = 0; j < 4; j++)
                if(((insm >> j) & 1u) && ground_col_needed(p, ixs[j]))
                {
                  atomicOr(gbev + (unsigned)iys[j] * (unsigned)p.wpr + (unsigned)(ixs[j] >> 5), 1u << (ixs[j] & 31));
                  W.rmin = min(W.rmin, iys[j]);
                  W.rmax = max(W.rmax, iys[j]);
                }
            }
          }
        }
        if(und)
        {
          labs[widx] = lw;
          word_prefetch_l2(FR, wbase + widx);
        }
      }
      if(__any_sync(0xffffffffu, und != 0u))
        n = compact_append(act, n, und, it, lane);
    }
    if(n == 0)
      continue;
    __syncwarp();
    qr_dense<SRC>(p, FR, S, W, F, amask, ground, gbev, wbase, n, warp, lane);
    __syncwarp();
  }
  if(run_n)
  {
    atomicAdd(&Q.sd[run_l], run_d);
    atomicAdd(&Q.sc[run_l], run_c);
    atomicAdd(&Q.sn[run_l], run_n);
  }
  qr_epilogue(S, W, F, ground, tid, lane); // (contains the block barrier)
  if(tid < SSD_GPU_MAX_PLATEAUS && Q.sn[tid])
  {
    atomicAdd(reinterpret_cast<unsigned long long *>(&F.plat[tid].sum_d), (unsigned long long)(long long)Q.sd[tid]);
    atomicAdd(&F.plat[tid].sum_c, (unsigned long long)Q.sc[tid]);
    atomicAdd(&F.plat[tid].n_sum, Q.sn[tid]);
    atomicAdd(&F.plat[tid].n_in_quad, Q.sn[tid]);
  }
}
