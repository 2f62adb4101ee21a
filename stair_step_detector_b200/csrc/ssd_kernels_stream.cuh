// ssd_kernels_stream.cuh -- the resident-frame chain (EXPERIMENTAL, SSD_GPU_PATH=resident): ONE persistent kernel reads every
// vertex exactly once. Parity-tested like the default chain; measured SLOWER (83 k frames/s against 268 k, DESIGN.md section 6,
// "Round 2"): it is kept as the record of that experiment and because k_quad_sum / the phase bodies are shared with the record chain.
//
// Why: labels depend on the whole-frame height histogram (pointcloud.cpp:184-256 before :280-343), so the classic chain
// (ssd_kernels_points.cuh) reads the plateau points of a frame three times from HBM (24 B/point of DRAM traffic against
// 13 algorithmic). A 1024x768 frame is 9.4 MB of vertices -- the shared memory of the 148 SMs together holds three of them.
// k_frame_stream keeps what the later phases need across the barrier the histogram imposes:
//
//   phase 1 (per 128-point sub-step, one warp): TMA bulk copy of the step's vertices into the warp's raw ring in shared memory
//            -> z>0 / CameraToWorld / range filter / height bin (the decisions of point_code_scaled, exact fallback: the same as
//            k_transform_bin, pointcloud.cpp:122-178) -> block histogram in shared memory (pointcloud.cpp:194-204) ->
//            a 4-byte record per in-range point {BEV pixel (pointcloud.cpp:79-83), parity of the bin, height offset inside the
//            bin} + the 1-byte bin code into the warp's record ring (global memory, re-used every SSD_GPU_FS_LAG frames).
//   frame barrier: the last warp of a CTA to leave a frame adds the CTA's histogram to the frame's (global reductions, no
//            fence, no counter); the frame's OWNER warp (by frame index) polls the frame's histogram, sees it complete when the
//            bins add up to the frame's point count, evaluates peaks / plateau bands / bin->label LUT (peaks_warp,
//            pointcloud.cpp:214-256, 300-335, 402-418) and raises the frame's flag; one warp per CTA polls it and copies the
//            LUT into the CTA's shared memory.
//   phase 2 (same warp, same step, from its record ring): bin code -> segment label (1 B/point, the only per-point
//            output), BEV occupancy bit of every point of an outlined plateau (projectToBinaryImage, pointcloud.cpp:458-471),
//            and one 16-byte summary per 32 points {label, count, sums of bin codes and height offsets, BEV pixel box} for the
//            per-step mean (calcAverageZ, pointcloud.cpp:574-581): k_quad_sum adds whole summaries whose pixel box lies inside
//            the quadrilateral and re-reads only the points of the summaries an edge crosses.
//
// Every warp is its own little pipeline (TMA issue -> phase 1 -> phase 2, phase 2 trailing by up to SSD_GPU_FS_LAG frames); the
// steps of a frame are dealt to the warps by a rotation that changes with the frame (FsCur); the only block-level state are
// the per-frame accumulators; there is no grid-wide barrier. All CTAs must be co-resident (cooperative launch, one CTA per SM).
#pragma once
#include "ssd_kernels_points.cuh"

#ifndef SSD_FS_WARPS
#define SSD_FS_WARPS 16
#endif
#define SSD_FS_THREADS (SSD_FS_WARPS * 32)
#define SSD_FS_NB 16             // frames a CTA can have in flight (accumulator / LUT slots)
#define SSD_FS_STEP_PX 128       // points per sub-step: 32 lanes x 4
#ifndef SSD_FS_SUB
#define SSD_FS_SUB 2              // sub-steps per step (one bulk copy, one ring slot, one turn of a warp's loop)
#endif
#define SSD_FS_REC_BYTES 640     // records of one sub-step: 32 code words + 32 x 4 records

// One summary per 32 consecutive pixels (8 lanes x 4): the points of the ground / outlined plateaus among them.
struct __align__(16) GroupSum
{
  unsigned short ixmin, ixmax, iymin, iymax; // BEV pixel box of those points
  int ds;                                    // sum of their height offsets d (units of 2^-rec_zshift m, relative to the bin centre)
  unsigned short cs;                         // sum of their bin codes
  unsigned char label;                       // their (common) segment label
  unsigned char count;                       // how many; 0: none; 0xff: mixed labels or an uncertain pixel -> per point in k_quad_sum
};
#define SSD_GS_COMPLEX 0xffu

struct FsParams
{
  int n_frames;
  int d_raw, d_rec;   // ring depths per warp (slots): raw vertices in shared memory, records in global memory (L2)
  unsigned char *recs; // record ring: grid x SSD_FS_WARPS x d_rec slots of SSD_FS_REC_BYTES
  int flags;          // (none defined)
  unsigned *done;     // per frame: CTAs that have delivered their histogram (self-resetting)
  GroupSum *sums;     // n_frames x steps x 4
  unsigned long long *prof; // optional cycle counters (SSD_GPU_FS_PROF=1), else nullptr
  size_t prof_warp_off;
};

struct FsAcc // per in-flight frame (slot f % SSD_FS_NB) of one CTA
{
  unsigned hist[SSD_BINS_PAD];
  int rmin[SSD_GPU_MAX_PLATEAUS], rmax[SSD_GPU_MAX_PLATEAUS];
  unsigned p1_left, p2_left, n_exact, oob, n_def, pad[3];
};

struct FsPeaks // scratch of the warp that evaluates a frame's peaks (one at a time per CTA: lock)
{
  PeaksScratch K;
  unsigned lock, pad[3];
};

struct FsLut // the CTA's copies of the frames' bin -> label LUTs (slot f % SSD_FS_NB), filled by whichever warp sees the frame ready first
{
  unsigned short lut[SSD_FS_NB][SSD_BINS_PAD];
  int frame1[SSD_FS_NB];    // frame + 1 whose LUT the slot holds
  unsigned busy[SSD_FS_NB]; // a warp of the CTA is asking global memory about this slot
};

template<class SRC>
struct FsSrc;
template<>
struct FsSrc<SrcVertices>
{
  static constexpr int STEP_BYTES = SSD_FS_STEP_PX * 12;
};
template<>
struct FsSrc<SrcDepth>
{
  static constexpr int STEP_BYTES = SSD_FS_STEP_PX * 2;
};

__device__ __forceinline__ const void *fs_src_base(const SrcVertices &s) { return s.xyz; }
__device__ __forceinline__ const void *fs_src_base(const SrcDepth &s) { return s.z16; }

__host__ __device__ inline size_t fs_smem_bytes(int step_bytes, int d_raw, int d_rec)
{
  (void)d_rec;
  return (size_t)SSD_FS_WARPS * d_raw * step_bytes * SSD_FS_SUB + sizeof(FsLut) +
         sizeof(FsAcc) * SSD_FS_NB + sizeof(FsPeaks) + (size_t)SSD_FS_WARPS * d_raw * 8 + 128;
}

// point_code_scaled (ssd_device.cuh) that also hands out the point's position inside its height bin:
// dfrac = (t - 0.5) - floor(t) in [-0.5, 0.5), t = (wz - z_min) * hir, exact remainder of the f32 value the bin came from.
__device__ __forceinline__ unsigned point_code_scaled_d(const DevParams &p, float x, float y, float z, bool &uncertain, float &dfrac)
{
  const float m = max3abs_nan(x, y, z);
  const float eps = fmaf(p.E1s, m, p.E0s);
  float vx, vy;
  f2_unpack(f2_affine(p.sxy2, p.sbxy2, x, y, z), vx, vy);
  const float vz = fmaf(p.sa[8], z, fmaf(p.sa[7], y, fmaf(p.sa[6], x, p.sb[2])));
  const float e1 = max3abs_nan(vx, vy, vz) - 1.0f;
  const float MAGIC = 12582912.0f;
  const float uf = fmaf(vz, p.Gf, p.Gm);
  const float s = uf + MAGIC;
  const float d = uf - (s - MAGIC);
  const float thr = fmaf(-p.Gup, eps, p.thr0);
  const bool out = e1 > eps;
  const bool in_bin = e1 < -eps && fabsf(d) < thr;
  const bool valid = z > 0.f;
  uncertain = valid && !(out || in_bin);
  dfrac = d;
  const unsigned c = out ? SSD_CODE_OUT_OF_RANGE : ((unsigned)__float_as_int(s) & 0xffu);
  return valid ? c : SSD_CODE_INVALID;
}

// exact fallback: the code by the double chain, and the height offset relative to THAT bin
__device__ __noinline__ unsigned point_code_slow_d(const DevParams &p, float fx, float fy, float fz, int &d)
{
  const unsigned c = point_code(p, fx, fy, fz);
  d = 0;
  if(c < SSD_CODE_OUT_OF_RANGE)
  {
    const double t = (camera_to_world_z(p, fx, fy, fz) - p.z_min) * p.hir;
    d = (int)rint((t - (double)c - 0.5) * (double)p.rec_mf);
  }
  return c;
}

// exact BEV pixel of an in-range point, packed as a record's pixel part; iy = H when it falls outside the image
__device__ __noinline__ unsigned pixel_slow(const DevParams &p, float fx, float fy, float fz)
{
  double wx, wy;
  camera_to_world_xy(p, fx, fy, fz, wx, wy);
  int ix, iy;
  world_to_image(p, wx, wy, ix, iy);
  if(ix >= 0 && ix < p.W && iy >= 0 && iy < p.H)
    return ((unsigned)iy << p.rec_bx) | (unsigned)ix;
  return (unsigned)p.H << p.rec_bx;
}

__device__ __forceinline__ unsigned long long fs_gtime()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// per-frame time stamps (ns) when profiling: [0] first CTA arrival, [1] last, [2] LUT published, [3] first / [4] last phase-2 step
#define FS_PROF_HDR 16
#define FS_PROF_PER_FRAME 8

__device__ __forceinline__ uint4 ld_volatile_v4(const void *ptr)
{
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr) : "memory");
  return v;
}

// peaks of a frame inside k_frame_stream: the CTA's scratch under its lock, then the flag the other warps poll
__device__ inline void fs_peaks(const DevParams &p, FrameDev &F, FsPeaks &K, const uint4 h0, const uint4 h1, int lane)
{
  if(lane == 0)
  {
    unsigned spin = 0;
    while(atomicCAS(&K.lock, 0u, 1u) != 0u)
    {
      __nanosleep(32);
      if(++spin > (1u << 24))
        __trap();
    }
  }
  __syncwarp();
  __threadfence_block();
  peaks_warp(p, F, K.K, h0, h1, lane);
  // the plateau records and the LUT must be visible before the flag that announces them
  __threadfence();
  __syncwarp();
  if(lane == 0)
  {
    *(volatile unsigned *)&F.ready = 1u;
    __threadfence_block();
    atomicExch(&K.lock, 0u);
  }
}

// a warp leaves phase 1 of frame f: the last warp of the CTA adds the CTA's histogram to the frame's (global reductions, no
// fence, no counter: the frame's owner sees it complete when the bins add up to the frame's point count)
__device__ inline void fs_leave_p1(FrameDev *frames, FsAcc *acc, int f, int lane, unsigned long long *prof)
{
  FsAcc &A = acc[f & (SSD_FS_NB - 1)];
  __syncwarp();
  unsigned last = 0;
  if(lane == 0)
  {
    __threadfence_block();
    last = atomicAdd(&A.p1_left, 1u) == SSD_FS_WARPS - 1 ? 1u : 0u;
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if(!last)
    return;
  __threadfence_block();
  volatile unsigned *vh = A.hist;
  for(int b = lane; b < SSD_BINS_PAD; b += 32)
  {
    const unsigned v = vh[b];
    if(v)
    {
      atomicAdd(&frames[f].hist[b], v);
      vh[b] = 0u;
    }
  }
  if(lane == 0)
  {
    const unsigned ne = *(volatile unsigned *)&A.n_exact;
    if(ne)
    {
      atomicAdd(&frames[f].n_exact_bin, ne);
      A.n_exact = 0;
    }
    *(volatile unsigned *)&A.p1_left = 0u;
    if(prof)
    {
      const unsigned long long t = fs_gtime();
      atomicMin(prof + FS_PROF_HDR + (size_t)f * FS_PROF_PER_FRAME + 0, t);
      atomicMax(prof + FS_PROF_HDR + (size_t)f * FS_PROF_PER_FRAME + 5, t);
    }
  }
}

// a warp leaves phase 2 of frame f: the last warp of the CTA delivers the CTA's per-plateau BEV row ranges and counters
__device__ inline void fs_leave_p2(FrameDev *frames, FsAcc *acc, int f, int lane)
{
  FsAcc &A = acc[f & (SSD_FS_NB - 1)];
  __syncwarp();
  unsigned last = 0;
  if(lane == 0)
  {
    __threadfence_block();
    last = atomicAdd(&A.p2_left, 1u) == SSD_FS_WARPS - 1 ? 1u : 0u;
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if(!last)
    return;
  __threadfence_block();
  FrameDev &F = frames[f];
  {
    volatile int *rmin = A.rmin, *rmax = A.rmax;
    const int hi = rmax[lane];
    if(hi >= 0)
    {
      atomicMin(&F.plat[lane].row_min, rmin[lane]);
      atomicMax(&F.plat[lane].row_max, hi);
      rmin[lane] = 0x7fffffff;
      rmax[lane] = -1;
    }
  }
  if(lane == 0)
  {
    if(*(volatile unsigned *)&A.oob)
    {
      atomicOr(&F.status, SSD_STATUS_BEV_OOB);
      A.oob = 0;
    }
    const unsigned nd = *(volatile unsigned *)&A.n_def;
    if(nd)
    {
      atomicAdd(&F.n_def_bev, nd);
      A.n_def = 0;
    }
    *(volatile unsigned *)&A.p2_left = 0u;
  }
  __threadfence_block();
}

// ---- phase 1 of one step: raw slot -> codes, histogram, records ----
__device__ __forceinline__ void fs_unpack_raw(const SrcVertices &, const DevParams &, const unsigned char *raw, unsigned, int lane, float vx[4],
                                              float vy[4], float vz[4])
{
  const float4 *s4 = reinterpret_cast<const float4 *>(raw) + lane * 3;
  const float4 v0 = s4[0], v1 = s4[1], v2 = s4[2];
  vx[0] = v0.x, vy[0] = v0.y, vz[0] = v0.z;
  vx[1] = v0.w, vy[1] = v1.x, vz[1] = v1.y;
  vx[2] = v1.z, vy[2] = v1.w, vz[2] = v2.x;
  vx[3] = v2.y, vy[3] = v2.z, vz[3] = v2.w;
}
__device__ __forceinline__ void fs_unpack_raw(const SrcDepth &src, const DevParams &p, const unsigned char *raw, unsigned step, int lane, float vx[4],
                                              float vy[4], float vz[4])
{
  FrameD f;
  f.d2 = nullptr;
  f.xn = src.xn;
  f.yn = src.yn;
  f.unit = src.unit;
  f.wmagic = src.wmagic;
  f.W = (unsigned)p.W;
  WordD w;
  w.d = reinterpret_cast<const uint2 *>(raw)[lane];
  const unsigned i0 = step * SSD_FS_STEP_PX + (unsigned)lane * 4u;
  const unsigned v = (unsigned)(((unsigned long long)i0 * f.wmagic) >> 40), u = i0 - v * f.W;
  w.x4 = __ldg(reinterpret_cast<const float4 *>(f.xn + u));
  w.y = __ldg(f.yn + v);
  word_unpack(f, w, vx, vy, vz);
}

// One 128-point sub-step. Returns through `A` / `rec`. Two speeds:
//   * every point of the warp's 128 is invalid or certainly outside the measuring range in x or y (the image rows that look
//     past the staircase: about four sub-steps in ten): only the packed x/y rows of the transform are evaluated; codes 254 / 255,
//     the histogram, no records;
//   * otherwise the full decision of point_code_scaled for all four points of every lane (no divergence), the exact fallback
//     for the uncertain ones, and the records of the in-range points.
template<class SRC, class ACC>
__device__ __forceinline__ void fs_phase1(const DevParams &p, const SRC &src, const unsigned char *raw, unsigned *code_slot, uint4 *rec_slot, ACC &A,
                                          unsigned step, int lane)
{
  float vx[4], vy[4], vz[4];
  fs_unpack_raw(src, p, raw, step, lane, vx, vy, vz);
  const float MAGIC = 12582912.0f;
  const int zsh = 32 - p.rec_zbits;
  float eps[4], sx[4], sy[4];
  bool skip = true;
#pragma unroll
  for(int j = 0; j < 4; j++)
  {
    const float m = max3abs_nan(vx[j], vy[j], vz[j]);
    eps[j] = fmaf(p.E1s, m, p.E0s);
    f2_unpack(f2_affine(p.sxy2, p.sbxy2, vx[j], vy[j], vz[j]), sx[j], sy[j]);
    // certainly out of range in x or y (then max|v_i| - 1 > eps as well: point_code_scaled says 254), or invalid (255)
    const bool far = fmaxf(fabsf(sx[j]), fabsf(sy[j])) - 1.0f > eps[j];
    skip = skip && (far || !(vz[j] > 0.f));
  }
  unsigned c[4];
  if(__all_sync(0xffffffffu, skip))
  {
#pragma unroll
    for(int j = 0; j < 4; j++)
    {
      c[j] = vz[j] > 0.f ? SSD_CODE_OUT_OF_RANGE : SSD_CODE_INVALID;
      atomicAdd(A.hist + c[j], 1u);
    }
    __stcg(code_slot, c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24));
    return;
  }
  unsigned zp[4];
  bool unc[4];
#pragma unroll
  for(int j = 0; j < 4; j++)
  {
    // the rest of point_code_scaled (ssd_device.cuh): z row, range test on all three rows, height bin with its certainty
    const float sz = fmaf(p.sa[8], vz[j], fmaf(p.sa[7], vy[j], fmaf(p.sa[6], vx[j], p.sb[2])));
    const float e1 = max3abs_nan(sx[j], sy[j], sz) - 1.0f;
    const float uf = fmaf(sz, p.Gf, p.Gm);
    const float s = uf + MAGIC;
    const float d = uf - (s - MAGIC);
    const float thr = fmaf(-p.Gup, eps[j], p.thr0);
    const bool out = e1 > eps[j];
    const bool in_bin = e1 < -eps[j] && fabsf(d) < thr;
    const bool valid = vz[j] > 0.f;
    unc[j] = valid && !(out || in_bin);
    const unsigned cc = out ? SSD_CODE_OUT_OF_RANGE : ((unsigned)__float_as_int(s) & 0xffu);
    c[j] = valid ? cc : SSD_CODE_INVALID;
    zp[j] = (unsigned)__float_as_int(fmaf(d, p.rec_mf, MAGIC)) << zsh;
  }
  if(unc[0] || unc[1] || unc[2] || unc[3])
  {
    unsigned ne = 0;
#pragma unroll
    for(int j = 0; j < 4; j++)
      if(unc[j])
      {
        int d;
        c[j] = point_code_slow_d(p, vx[j], vy[j], vz[j], d);
        zp[j] = (unsigned)d << zsh;
        ne++;
      }
    atomicAdd(&A.n_exact, ne);
  }
  const unsigned cw = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
#pragma unroll
  for(int j = 0; j < 4; j++)
    atomicAdd(A.hist + c[j], 1u);
  __stcg(code_slot, cw);
  // records of the lane's in-range points (codes 254 / 255: out of range / invalid)
  if((cw & 0xfefefefeu) != 0xfefefefeu)
  {
    unsigned r[4], slow = 0;
#pragma unroll
    for(int j = 0; j < 4; j++)
    {
      int ix, iy;
      const bool ok = fast_pixel2(p, vx[j], vy[j], vz[j], ix, iy);
      r[j] = zp[j] | ((c[j] & 1u) << (p.rec_bx + p.rec_by)) | ((unsigned)iy << p.rec_bx) | (unsigned)ix;
      slow |= (!ok && c[j] < SSD_CODE_OUT_OF_RANGE) ? (1u << j) : 0u;
    }
    if(slow)
    {
      // The single-precision pixel of an in-range point was not certain (about one point in 200): the exact double chain
      // decides here, where the vertex is in registers. A pixel outside the image (the x == W wrap of pointcloud.cpp:81,468)
      // is marked iy = H: phase 2 then follows the reference's unchecked write from the vertex itself.
      unsigned ns = 0;
#pragma unroll
      for(int j = 0; j < 4; j++)
        if((slow >> j) & 1u)
        {
          r[j] = zp[j] | ((c[j] & 1u) << (p.rec_bx + p.rec_by)) | pixel_slow(p, vx[j], vy[j], vz[j]);
          ns++;
        }
      atomicAdd(&A.n_def, ns);
    }
    __stcg(rec_slot, make_uint4(r[0], r[1], r[2], r[3]));
  }
}

// ---- phase 2 of one step: record slot -> labels, BEV bits, summaries ----
template<class SRC, class ACC>
__device__ __forceinline__ void fs_phase2(const DevParams &p, const SRC &src, const FsParams &a, const unsigned cw, const uint4 rv, const unsigned short *lut,
                                          ACC &A, unsigned frame, unsigned step, unsigned char *labels, unsigned *bev, unsigned bmw, int lane)
{
  const size_t word = (size_t)frame * (size_t)(p.N >> 2) + (size_t)step * 32u + (unsigned)lane;
  GroupSum *gs = a.sums + (((size_t)frame * (size_t)p.gs_steps + step) * 4u + (unsigned)(lane >> 3));
  unsigned lab = cw, fl = 0, ol = 0;
  if((cw & 0xfefefefeu) != 0xfefefefeu)
  {
    const unsigned e0 = lut[cw & 0xffu], e1 = lut[(cw >> 8) & 0xffu], e2 = lut[(cw >> 16) & 0xffu], e3 = lut[cw >> 24];
    lab = (e0 & 0xffu) | ((e1 & 0xffu) << 8) | ((e2 & 0xffu) << 16) | (e3 << 24);
    const unsigned K = SSD_LUT_OUTLINED | SSD_LUT_GROUND;
    fl = ((e0 & K) ? 1u : 0u) | ((e1 & K) ? 2u : 0u) | ((e2 & K) ? 4u : 0u) | ((e3 & K) ? 8u : 0u);
    ol = ((e0 >> 8) & 1u) | ((e1 >> 7) & 2u) | ((e2 >> 6) & 4u) | ((e3 >> 5) & 8u);
  }
  reinterpret_cast<unsigned *>(labels)[word] = lab;
  if(!__any_sync(0xffffffffu, fl != 0u))
  {
    if((lane & 7) == 0)
      *reinterpret_cast<uint4 *>(gs) = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const int bx = p.rec_bx, zsh = 32 - p.rec_zbits;
  const unsigned mx = (1u << bx) - 1u, my = (1u << p.rec_by) - 1u;
  int ixmin = 0x7fff, ixmax = 0, iymin = 0x7fff, iymax = 0, dsum = 0;
  unsigned csum = 0, bad = 0;
  // label of the lane's first flagged point; the lane is "uniform" when all its flagged points carry it
  const unsigned l0 = (lab >> (8 * (__ffs(fl | 16u) - 1) & 31)) & 0xffu;
  if(fl)
  {
    const unsigned r[4] = { rv.x, rv.y, rv.z, rv.w };
    const unsigned bytes = ((fl * 0x00204081u) & 0x01010101u) * 0xffu;
    bad = ((lab ^ (l0 * 0x01010101u)) & bytes) != 0u;
    unsigned lidx = (frame * SSD_GPU_MAX_PLATEAUS + l0) * bmw;
#pragma unroll
    for(int j = 0; j < 4; j++)
      if((fl >> j) & 1u)
      {
        const int ix = (int)(r[j] & mx), iy = (int)((r[j] >> bx) & my);
        const unsigned l = (lab >> (8 * j)) & 0xffu;
        if(iy >= p.H)
        {
          // the single-precision pixel was not certain (fast_pixel2): the exact double chain decides, from the vertex itself
          bad = 1;
          if((ol >> j) & 1u)
          {
            const typename SrcTraits<SRC>::Frame FR = src_frame(src, p, (size_t)frame * p.N);
            float fx, fy, fz;
            point_load(FR, step * SSD_FS_STEP_PX + (unsigned)lane * 4u + (unsigned)j, fx, fy, fz);
            double wx, wy;
            camera_to_world_xy(p, fx, fy, fz, wx, wy);
            int x, y;
            if(bev_pixel(p, wx, wy, x, y))
            {
              atomicOr(bev + ((size_t)(frame * SSD_GPU_MAX_PLATEAUS + l) * bmw + (size_t)y * p.wpr + (x >> 5)), 1u << (x & 31));
              atomicMin(&A.rmin[l], y);
              atomicMax(&A.rmax[l], y);
            }
            else
              A.oob = 1;
          }
          continue;
        }
        if((ol >> j) & 1u)
        {
          const unsigned base = bad ? (frame * SSD_GPU_MAX_PLATEAUS + l) * bmw : lidx;
          atomicOr(bev + (base + (unsigned)iy * (unsigned)p.wpr + ((unsigned)ix >> 5)), 1u << (ix & 31));
        }
        ixmin = min(ixmin, ix);
        ixmax = max(ixmax, ix);
        iymin = min(iymin, iy);
        iymax = max(iymax, iy);
        dsum += (int)r[j] >> zsh;
        csum += (cw >> (8 * j)) & 0xffu;
      }
  }
  // 8-lane summary by xor butterflies over packed words (a reduction instruction with a partial member mask would be executed
  // once per group, serialised): {ixmin, iymin} / {ixmax, iymax} as 16-bit pairs (two-lane min / max instructions),
  // {count, sum of codes} and the sum of the height offsets as plain sums. Uniformity of the label by ballots.
  const unsigned gm = 0xffu << (lane & 24);
  const unsigned fg = __ballot_sync(0xffffffffu, fl != 0u) & gm;
  const unsigned ll = __shfl_sync(0xffffffffu, l0, fg ? __ffs(fg) - 1 : lane); // label of the group's first flagged lane
  const unsigned nb = __ballot_sync(0xffffffffu, bad || (fl && l0 != ll)) & gm;
  const unsigned gol = __ballot_sync(0xffffffffu, ol != 0u) & gm;
  unsigned gmn = (unsigned)ixmin | ((unsigned)iymin << 16), gmx = (unsigned)ixmax | ((unsigned)iymax << 16);
  unsigned cc = (unsigned)__popc(fl) | (csum << 8);
  int gds = dsum;
#pragma unroll
  for(int k = 1; k < 8; k <<= 1)
  {
    gmn = __vminu2(gmn, __shfl_xor_sync(0xffffffffu, gmn, k));
    gmx = __vmaxu2(gmx, __shfl_xor_sync(0xffffffffu, gmx, k));
    cc += __shfl_xor_sync(0xffffffffu, cc, k);
    gds += __shfl_xor_sync(0xffffffffu, gds, k);
  }
  const unsigned gcnt = cc & 0xffu;
  const bool uniform = nb == 0u && fg != 0u;
  const int gymin = (int)(gmn >> 16), gymax = (int)(gmx >> 16);
  const unsigned lmin = ll;
  if((lane & 7) == 0)
  {
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if(gcnt)
    {
      if(uniform)
      {
        o.x = (gmn & 0xffffu) | (gmx << 16);
        o.y = (gmn >> 16) | (gmx & 0xffff0000u);
        o.z = (unsigned)gds;
        o.w = (cc >> 8) | (lmin << 16) | (gcnt << 24);
      }
      else
        o.w = SSD_GS_COMPLEX << 24;
    }
    *reinterpret_cast<uint4 *>(gs) = o;
  }
  // BEV rows touched, per plateau (the outline stages only that band): one pair of shared-memory reductions per group
  // (all flagged points of a uniform group share the label, so any lane's `ol` tells whether it is outlined)
  if(uniform && gol)
  {
    if((lane & 7) == 0 && gymax >= gymin && gcnt)
    {
      atomicMin(&A.rmin[lmin], gymin);
      atomicMax(&A.rmax[lmin], gymax);
    }
  }
  else if(ol && iymax >= iymin)
  {
    // mixed labels inside the group (plateau boundaries in the image): per point
    const unsigned r[4] = { rv.x, rv.y, rv.z, rv.w };
#pragma unroll
    for(int j = 0; j < 4; j++)
      if((ol >> j) & 1u)
      {
        const int iy = (int)((r[j] >> bx) & my);
        if(iy < p.H)
        {
          const unsigned l = (lab >> (8 * j)) & 0xffu;
          atomicMin(&A.rmin[l], iy);
          atomicMax(&A.rmax[l], iy);
        }
      }
  }
}

// ---------------------------------------------------------------------------------------------
// k_frame_stream. grid = one CTA per SM (cooperative launch: all CTAs co-resident), block = 16 warps.
// ---------------------------------------------------------------------------------------------
// A warp's position in its step sequence. In frame f it takes the steps s_j = j * TW + ((warp - f * ROT + j * R) mod TW), j = 0, 1, ...
// while s_j < steps per frame (TW = warps of the grid): every block of TW steps is a rotation of the warps, so the steps of a
// frame are dealt out exactly once; the rotations by frame (ROT) and by block (R) move a warp across the image columns and
// rows from step to step -- with a fixed stride a warp would stay in one column block of the image for ever, and the warps
// looking at the staircase would do three times the work of the ones looking past it (measured).
struct FsCur
{
  unsigned k, f, s, j, base;
};
#define SSD_FS_ROT 5u
#define SSD_FS_R 3u

template<class SRC>
__global__ void __launch_bounds__(SSD_FS_THREADS, 1) k_frame_stream(const __grid_constant__ DevParams p, const SRC src, const __grid_constant__ FsParams a,
                                                                    unsigned char *__restrict__ labels, FrameDev *__restrict__ frames,
                                                                    unsigned *__restrict__ bev, size_t bm_words)
{
  extern __shared__ __align__(128) unsigned char fs_smem[];
  constexpr int SUBRAW = FsSrc<SRC>::STEP_BYTES;
  constexpr int RAW = SUBRAW * SSD_FS_SUB;
  constexpr int REC = SSD_FS_REC_BYTES * SSD_FS_SUB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d_raw = a.d_raw, d_rec = a.d_rec;
  unsigned char *raw_all = fs_smem;
  FsLut *lutc = reinterpret_cast<FsLut *>(raw_all + (size_t)SSD_FS_WARPS * d_raw * RAW);
  FsAcc *acc = reinterpret_cast<FsAcc *>(lutc + 1);
  FsPeaks *pk = reinterpret_cast<FsPeaks *>(acc + SSD_FS_NB);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(pk + 1);

  {
    unsigned *z = reinterpret_cast<unsigned *>(acc);
    for(int i = tid; i < (int)(sizeof(FsAcc) * SSD_FS_NB / 4); i += SSD_FS_THREADS)
      z[i] = 0u;
    __syncthreads();
    for(int i = tid; i < SSD_FS_NB * SSD_GPU_MAX_PLATEAUS; i += SSD_FS_THREADS)
    {
      acc[i / SSD_GPU_MAX_PLATEAUS].rmin[i % SSD_GPU_MAX_PLATEAUS] = 0x7fffffff;
      acc[i / SSD_GPU_MAX_PLATEAUS].rmax[i % SSD_GPU_MAX_PLATEAUS] = -1;
    }
    if(tid == 0)
      pk->lock = 0u;
    if(tid < SSD_FS_NB)
    {
      lutc->frame1[tid] = 0;
      lutc->busy[tid] = 0u;
    }
    if(lane == 0)
    {
      for(int i = 0; i < d_raw; i++)
        mbar_init((unsigned)__cvta_generic_to_shared(bars + warp * d_raw + i), 1);
      mbar_fence_init();
    }
    __syncthreads();
  }

  unsigned char *raw = raw_all + (size_t)warp * d_raw * RAW;
  unsigned char *rec = a.recs + ((size_t)blockIdx.x * SSD_FS_WARPS + warp) * (size_t)d_rec * REC;
  const unsigned raw_sa = (unsigned)__cvta_generic_to_shared(raw);
  const unsigned bar_sa = (unsigned)__cvta_generic_to_shared(bars + warp * d_raw);

  const unsigned S = (unsigned)p.gs_steps / SSD_FS_SUB; // steps per frame
  const unsigned TW = gridDim.x * SSD_FS_WARPS, gw = blockIdx.x * SSD_FS_WARPS + warp;
  const unsigned n_frames = (unsigned)a.n_frames;
  FsCur ct, c1, c2;
  // (a warp whose rotated position lies past the frame's last step has no step in that frame)
#define FS_NEXT_FRAME(c)                                                                        \
  do                                                                                            \
  {                                                                                             \
    (c).f++;                                                                                    \
    (c).j = 0;                                                                                  \
    (c).base = (c).base >= SSD_FS_ROT ? (c).base - SSD_FS_ROT : (c).base + TW - SSD_FS_ROT;     \
    (c).s = (c).base;                                                                           \
  } while((c).s >= S && (c).f < n_frames)
#define FS_ADV(c)                                                        \
  do                                                                     \
  {                                                                      \
    (c).k++;                                                             \
    (c).j++;                                                             \
    unsigned b_ = (c).base + (c).j * SSD_FS_R;                           \
    b_ = b_ >= TW ? b_ - TW : b_;                                        \
    (c).s = (c).j * TW + b_;                                             \
    if((c).s >= S)                                                       \
      FS_NEXT_FRAME(c);                                                  \
  } while(0)
  ct.k = 0, ct.f = 0, ct.j = 0, ct.base = gw, ct.s = gw;
  if(ct.s >= S)
    FS_NEXT_FRAME(ct);
  c1 = ct;
  c2 = ct;
#define FS_LIVE(c) ((c).f < n_frames)
#define FS_FRAME(c) ((c).f < n_frames ? (c).f : n_frames)
  unsigned slot_t = 0, slot_1 = 0, phase_1 = 0, rslot_1 = 0, rslot_2 = 0;
  unsigned left1 = 0, left2 = 0; // frames this warp has left (phase 1 / phase 2): [0, left)
  int lut_f = -1;                // frame whose LUT the warp's copy holds
  const unsigned bmw = (unsigned)bm_words;
  const unsigned char *gsrc = reinterpret_cast<const unsigned char *>(fs_src_base(src));

  // frames before the warp's first step
  for(; left1 < FS_FRAME(c1); left1++)
    fs_leave_p1(frames, acc, (int)left1, lane, a.prof);
  for(; left2 < FS_FRAME(c2); left2++)
    fs_leave_p2(frames, acc, (int)left2, lane);

  // frames whose peaks this warp evaluates: CTA f % grid, warp (f / grid) % 16
  unsigned own_f = blockIdx.x + gridDim.x * (unsigned)warp;
  const unsigned own_stride = gridDim.x * SSD_FS_WARPS;
  unsigned idle = 0;
  long long t_leave = 0, t_other = 0;
  long long t_p1 = 0, t_p2 = 0, t_wait = 0, n_idle = 0, n_poll = 0, n_i_done = 0, n_i_ring = 0, n_i_guard = 0, lag_sum = 0;
  const long long t_begin = clock64();
  const bool prof = a.prof != nullptr;
  unsigned pf_k = 0xffffffffu; // step whose records sit in (cw2, rv2)
  unsigned cw2 = 0;
  uint4 rv2 = make_uint4(0u, 0u, 0u, 0u);
  long long t_it_did = 0, t_it_idle = 0, t_sec_a = 0, t_sec_b = 0;
  while(FS_LIVE(c2) || own_f < n_frames)
  {
    const long long ti0 = prof ? clock64() : 0;
    // 1. keep the raw ring full
    while(FS_LIVE(ct) && ct.k - c1.k < (unsigned)d_raw)
    {
      if(lane == 0)
      {
        const unsigned bar = bar_sa + slot_t * 8u;
        mbar_expect_tx(bar, (unsigned)RAW);
        bulk_g2s(raw_sa + slot_t * (unsigned)RAW, gsrc + ((size_t)ct.f * S + ct.s) * (size_t)RAW, (unsigned)RAW, bar);
      }
      slot_t = slot_t + 1 == (unsigned)d_raw ? 0u : slot_t + 1;
      FS_ADV(ct);
    }
    const long long ti1 = prof ? clock64() : 0;
    t_sec_a += ti1 - ti0;
    const bool pending = c2.k < c1.k;
    // 2. the records of the oldest pending phase-2 step: requested now (from L2, where phase 1 left them), used after phase 1
    if(pending && pf_k != c2.k)
    {
      const unsigned char *r2 = rec + (size_t)rslot_2 * REC;
      cw2 = __ldcg(reinterpret_cast<const unsigned *>(r2) + lane);
      rv2 = __ldcg(reinterpret_cast<const uint4 *>(r2 + 128) + lane);
      pf_k = c2.k;
    }
    // 3. is that step's frame ready? The CTA keeps one copy of each frame's LUT; one warp at a time asks global memory
    //    (a 4-byte flag: 148 pollers, not 2368), asynchronously: the answer is looked at after phase 1
    const unsigned ls = c2.f & (SSD_FS_NB - 1);
    bool polled = false;
    unsigned pv = 0;
    if(pending && (int)c2.f != lut_f)
    {
      if(*(volatile int *)&lutc->frame1[ls] == (int)c2.f + 1)
      {
        __threadfence_block();
        lut_f = (int)c2.f;
      }
      else
      {
        unsigned got = 0;
        if(lane == 0)
          got = atomicCAS(&lutc->busy[ls], 0u, 1u) == 0u ? 1u : 0u;
        got = __shfl_sync(0xffffffffu, got, 0);
        if(got)
        {
          if(lane == 0)
            asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(pv) : "l"(&frames[c2.f].ready) : "memory");
          polled = true;
          n_poll++;
        }
      }
    }
    // 3b. the frame this warp owns: is its histogram complete? (asynchronous as well)
    const bool own_poll = own_f < n_frames && own_f < left1;
    uint4 oh0 = make_uint4(0u, 0u, 0u, 0u), oh1 = oh0;
    if(own_poll)
    {
      const uint4 *hsrc = reinterpret_cast<const uint4 *>(frames[own_f].hist) + lane * 2;
      oh0 = ld_volatile_v4(hsrc);
      oh1 = ld_volatile_v4(hsrc + 1);
    }
    bool did = false;
    if(prof)
      t_sec_b += clock64() - ti1;
    // 4. phase 1 of the next loaded step
    if(FS_LIVE(c1) && c1.k - c2.k < (unsigned)d_rec && c1.f - c2.f < (unsigned)(SSD_FS_NB - 2))
    {
      const long long tw0 = prof ? clock64() : 0;
      mbar_wait(bar_sa + slot_1 * 8u, phase_1);
      const long long t0 = prof ? clock64() : 0;
      t_wait += t0 - tw0;
#pragma unroll 1
      for(int i = 0; i < SSD_FS_SUB; i++)
      {
        unsigned char *ro = rec + (size_t)rslot_1 * REC + i * SSD_FS_REC_BYTES;
        fs_phase1(p, src, raw + (size_t)slot_1 * RAW + i * SUBRAW, reinterpret_cast<unsigned *>(ro) + lane, reinterpret_cast<uint4 *>(ro + 128) + lane,
                  acc[c1.f & (SSD_FS_NB - 1)], c1.s * SSD_FS_SUB + i, lane);
      }
      __syncwarp();
      if(prof)
        t_p1 += clock64() - t0;
      if(++slot_1 == (unsigned)d_raw)
      {
        slot_1 = 0;
        phase_1 ^= 1u;
      }
      rslot_1 = rslot_1 + 1 == (unsigned)d_rec ? 0u : rslot_1 + 1;
      FS_ADV(c1);
      {
        const long long tl = prof ? clock64() : 0;
        for(; left1 < FS_FRAME(c1); left1++)
          fs_leave_p1(frames, acc, (int)left1, lane, a.prof);
        if(prof)
          t_leave += clock64() - tl;
      }
      did = true;
    }
    // 5. the poll's answer: copy the frame's LUT into the CTA's slot
    if(polled)
    {
      pv = __shfl_sync(0xffffffffu, pv, 0);
      if(pv)
      {
        __threadfence();
        const uint4 v = ld_volatile_v4(reinterpret_cast<const uint4 *>(frames[c2.f].lut16) + lane);
        reinterpret_cast<uint4 *>(lutc->lut[ls])[lane] = v;
        __threadfence_block();
        __syncwarp();
        if(lane == 0)
          *(volatile int *)&lutc->frame1[ls] = (int)c2.f + 1;
        lut_f = (int)c2.f;
      }
      if(lane == 0)
      {
        __threadfence_block();
        atomicExch(&lutc->busy[ls], 0u);
      }
    }
    // 5b. the owned frame: every point counted -> peaks, plateau bands, LUT (the bins are final: each is written by
    //     atomic additions only and the total is the frame's point count)
    if(own_poll)
    {
      const unsigned tot = __reduce_add_sync(0xffffffffu, oh0.x + oh0.y + oh0.z + oh0.w + oh1.x + oh1.y + oh1.z + oh1.w);
      if(tot == (unsigned)p.N)
      {
        if(prof && lane == 0)
          a.prof[FS_PROF_HDR + (size_t)own_f * FS_PROF_PER_FRAME + 1] = fs_gtime();
        const long long tk = clock64();
        fs_peaks(p, frames[own_f], *pk, oh0, oh1, lane);
        t_other += clock64() - tk;
        if(prof && lane == 0)
          a.prof[FS_PROF_HDR + (size_t)own_f * FS_PROF_PER_FRAME + 2] = fs_gtime();
        own_f += own_stride;
        did = true;
      }
    }
    // 6. phase 2 of the oldest pending step
    if(pending && (int)c2.f == lut_f)
    {
      const long long t0 = prof ? clock64() : 0;
#pragma unroll 1
      for(int i = 0; i < SSD_FS_SUB; i++)
      {
        // the next sub-step's records are requested before this one is worked on
        unsigned cwn = 0;
        uint4 rvn = make_uint4(0u, 0u, 0u, 0u);
        if(i + 1 < SSD_FS_SUB)
        {
          const unsigned char *r2 = rec + (size_t)rslot_2 * REC + (i + 1) * SSD_FS_REC_BYTES;
          cwn = __ldcg(reinterpret_cast<const unsigned *>(r2) + lane);
          rvn = __ldcg(reinterpret_cast<const uint4 *>(r2 + 128) + lane);
        }
        fs_phase2(p, src, a, cw2, rv2, lutc->lut[ls], acc[ls], c2.f, c2.s * SSD_FS_SUB + i, labels, bev, bmw, lane);
        cw2 = cwn;
        rv2 = rvn;
      }
      __syncwarp();
      rslot_2 = rslot_2 + 1 == (unsigned)d_rec ? 0u : rslot_2 + 1;
      if(prof && lane == 0)
      {
        const unsigned long long t = fs_gtime();
        atomicMin(a.prof + FS_PROF_HDR + (size_t)c2.f * FS_PROF_PER_FRAME + 3, t);
        atomicMax(a.prof + FS_PROF_HDR + (size_t)c2.f * FS_PROF_PER_FRAME + 4, t);
      }
      FS_ADV(c2);
      for(; left2 < FS_FRAME(c2); left2++)
        fs_leave_p2(frames, acc, (int)left2, lane);
      if(prof)
        t_p2 += clock64() - t0;
      did = true;
    }
    if(did)
    {
      idle = 0;
      if(prof)
        t_it_did += clock64() - ti0;
    }
    else
    {
      // nothing to do: the ring is full behind a frame that is not ready. Back off (a spinning warp takes issue slots from
      // the working ones); a frame barrier that never completes (a CTA not resident?) traps instead of hanging the GPU
      n_idle++;
      if(prof)
      {
        if(!FS_LIVE(c1))
          n_i_done++;
        else if(c1.k - c2.k >= (unsigned)d_rec)
          n_i_ring++;
        else
          n_i_guard++;
        lag_sum += (long long)(c1.f - c2.f);
      }
      if(++idle > (1u << 21))
        __trap();
      __nanosleep(idle < 4 ? 250 : 1000);
      if(prof)
        t_it_idle += clock64() - ti0;
    }
  }
  if(prof && lane == 0)
  {
    atomicAdd(a.prof + 0, (unsigned long long)t_p1);
    atomicAdd(a.prof + 1, (unsigned long long)t_p2);
    atomicAdd(a.prof + 2, (unsigned long long)(clock64() - t_begin));
    atomicAdd(a.prof + 3, (unsigned long long)n_idle);
    atomicAdd(a.prof + 4, (unsigned long long)n_poll);
    atomicAdd(a.prof + 5, 1ull);
    atomicAdd(a.prof + 6, (unsigned long long)t_wait);
    if(a.n_frames >= 64)
    {
      // per-warp record behind the per-frame stamps (host: FS_PROF_HDR + FS_PROF_PER_FRAME * chunk_frames + 8 * global warp)
      unsigned long long *w = a.prof + a.prof_warp_off + (size_t)gw * 8;
      w[0] = (unsigned long long)t_p1, w[1] = (unsigned long long)t_p2, w[2] = (unsigned long long)t_wait, w[3] = (unsigned long long)n_idle;
      w[4] = (unsigned long long)(clock64() - t_begin), w[5] = (unsigned long long)n_i_ring, w[6] = (unsigned long long)t_leave, w[7] = (unsigned long long)t_other;
    }
    atomicAdd(a.prof + 11, (unsigned long long)t_it_did);
    atomicAdd(a.prof + 12, (unsigned long long)t_it_idle);
    atomicAdd(a.prof + 13, (unsigned long long)t_sec_a);
    atomicAdd(a.prof + 14, (unsigned long long)t_sec_b);
    atomicAdd(a.prof + 7, (unsigned long long)n_i_done);
    atomicAdd(a.prof + 8, (unsigned long long)n_i_ring);
    atomicAdd(a.prof + 9, (unsigned long long)n_i_guard);
    atomicAdd(a.prof + 10, (unsigned long long)lag_sum);
  }
#undef FS_ADV
#undef FS_NEXT_FRAME
#undef FS_FRAME
#undef FS_LIVE
}

// ---------------------------------------------------------------------------------------------
// k_quad_sum: getPointsInQuadrilateral + calcAverageZ (pointcloud.cpp:560-581) from the summaries k_frame_stream left.
// One summary (32 pixels) per lane and warp-tile. A summary whose BEV pixel box -- widened to the world rectangle every
// point with such a pixel can lie in -- sits inside the verified inner box of its step's QuadrilateralTest
// (quadtest_inner_box: isPointWithin() is true on the whole box) contributes its count and height sums wholesale; one
// entirely beyond the reject box contributes nothing; the rest (an edge of the quadrilateral crosses it, mixed labels, an
// uncertain pixel, ground points whose BEV columns detectFrontEdge looks at) goes point by point through the same dense
// pass as k_quad_reduce (qr_dense), re-reading those 32 vertices.
// ---------------------------------------------------------------------------------------------
struct QuadSumShared
{
  int sd[SSD_GPU_MAX_PLATEAUS];
  unsigned sc[SSD_GPU_MAX_PLATEAUS], sn[SSD_GPU_MAX_PLATEAUS];
};

template<class SRC>
__global__ void __launch_bounds__(SSD_PT_THREADS, SSD_QR_MINB) k_quad_sum(const __grid_constant__ DevParams p, const SRC src,
                                                                          const unsigned char *__restrict__ labels, FrameDev *__restrict__ frames,
                                                                          unsigned *__restrict__ bev, size_t bm_words, const GroupSum *__restrict__ sums,
                                                                          const uint4 *__restrict__ recs)
{
  __shared__ QuadReduceShared S;
  __shared__ QuadSumShared Q;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int frame = blockIdx.y;
  FrameDev &F = frames[frame];
  const unsigned amask = F.quad_amask;
  if(amask == 0u)
    return; // no valid plateau: nothing is emitted (pointcloud.cpp:434)
  const int ground = F.ground_index;
  const size_t fbase = (size_t)frame * p.N;
  const unsigned *lab32 = reinterpret_cast<const unsigned *>(labels + fbase);
  unsigned *gbev = ground >= 0 ? bev + ((size_t)frame * SSD_GPU_MAX_PLATEAUS + ground) * bm_words : nullptr;
  const int ngroups = p.N >> 5;
  const uint4 *gs = reinterpret_cast<const uint4 *>(sums) + (size_t)frame * ngroups;
  int wt, wt_end, wt_stride;
  warp_tile_range((p.N + SSD_WT_PX - 1) / SSD_WT_PX, warp, wt, wt_end, wt_stride);

  qr_init(S, F, amask, tid);
  if(tid < SSD_GPU_MAX_PLATEAUS)
  {
    Q.sd[tid] = 0;
    Q.sc[tid] = 0;
    Q.sn[tid] = 0;
  }
  __syncthreads();

  unsigned short *act = S.L.act[warp];
  unsigned *labs = S.L.lab[warp];
  QrWarp W = { 0xffu, 0u, 0u, 0u, 0ull, 0x7fffffff, -1 };
  const typename SrcTraits<SRC>::Frame FR = src_frame(src, p, fbase);
  // the lane's running run of whole summaries (consecutive summaries of a lane mostly share the label)
  unsigned run_l = 0xffu, run_n = 0, run_c = 0;
  int run_d = 0;
  const int c0 = p.W / 2 - 2;

  for(; wt < wt_end; wt += wt_stride)
  {
    const unsigned wbase = (unsigned)wt * (SSD_WT_PX / 4);
    const int g = wt * 32 + lane;
    uint4 gv = make_uint4(0u, 0u, 0u, 0u);
    if(g < ngroups)
      gv = __ldg(gs + g);
    if(wt + wt_stride < wt_end)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(gs + (size_t)(wt + wt_stride) * 32 + lane));
    const unsigned count = gv.w >> 24, l = (gv.w >> 16) & 0xffu;
    bool per_point = count == SSD_GS_COMPLEX, gbev_group = false;
    if(count != 0u && !per_point && ((amask >> (l & 31u)) & 1u) && l < SSD_GPU_MAX_PLATEAUS)
    {
      const int ixmin = (int)(gv.x & 0xffffu), ixmax = (int)(gv.x >> 16), iymin = (int)(gv.y & 0xffffu), iymax = (int)(gv.y >> 16);
      // world rectangle of the pixel box (f32; gs_margin covers the rounding of these four values and of the differences below)
      const float x0 = fmaf((float)ixmin, p.gs_xw, p.gs_x0), x1 = fmaf((float)(ixmax + 1), p.gs_xw, p.gs_x0);
      const float y1 = fmaf(-(float)iymin, p.gs_yw, p.gs_y0), y0 = fmaf(-(float)(iymax + 1), p.gs_yw, p.gs_y0);
      const float4 ib = S.fast[l].ibe;
      const float2 rj = S.fast[l].rj;
      const float ax = fmaxf(fabsf(x0 - ib.x), fabsf(x1 - ib.x)), ay = fmaxf(fabsf(y0 - ib.y), fabsf(y1 - ib.y));
      const bool inside = ax < ib.z - p.gs_margin && ay < ib.w - p.gs_margin;
      const bool outside = x0 - ib.x > rj.x + p.gs_margin || ib.x - x1 > rj.x + p.gs_margin || y0 - ib.y > rj.y + p.gs_margin ||
                           ib.y - y1 > rj.y + p.gs_margin;
      bool whole = inside;
      if(inside && (int)l == ground)
      {
        // ground points in the pixel columns detectFrontEdge probes need their BEV bit: from the records when the chain kept
        // them (every point of the summary is inside the quadrilateral), else point by point
        const unsigned r = (unsigned)(ixmin - c0 + 50 * 128) % 50u;
        if(r < 5u || r + (unsigned)(ixmax - ixmin) >= 50u)
        {
          if(recs)
            gbev_group = true;
          else
            whole = false;
        }
      }
      if(whole)
      {
        if(l != run_l)
        {
          if(run_n)
          {
            atomicAdd(&Q.sd[run_l], run_d);
            atomicAdd(&Q.sc[run_l], run_c);
            atomicAdd(&Q.sn[run_l], run_n);
          }
          run_l = l;
          run_d = 0;
          run_c = 0;
          run_n = 0;
        }
        run_d += (int)gv.z;
        run_c += gv.w & 0xffffu;
        run_n += count;
      }
      else if(!outside)
        per_point = true;
    }
    {
      // ground summaries taken whole: the BEV bits of their points in the probed columns, one pixel per lane
      unsigned gbm = __ballot_sync(0xffffffffu, gbev_group);
      const unsigned rmx = (1u << p.rec_bx) - 1u, rmy = (1u << p.rec_by) - 1u;
      while(gbm)
      {
        const int b = __ffs(gbm) - 1;
        gbm &= gbm - 1u;
        const size_t px = (size_t)(wt * 32 + b) * 32 + (unsigned)lane;
        const unsigned lb = __ldg(labels + fbase + px);
        const unsigned rc = __ldg(reinterpret_cast<const unsigned *>(recs + (size_t)frame * (p.N >> 2)) + px);
        const int ix = (int)(rc & rmx), iy = (int)((rc >> p.rec_bx) & rmy);
        if((int)lb == ground && ground_col_needed(p, ix))
        {
          atomicOr(gbev + (unsigned)iy * (unsigned)p.wpr + (unsigned)(ix >> 5), 1u << (ix & 31));
          W.rmin = min(W.rmin, iy);
          W.rmax = max(W.rmax, iy);
        }
      }
    }
    if(!__any_sync(0xffffffffu, per_point))
      continue;
    // ---- compaction of the 4-point words of the per-point summaries (word index within the warp-tile: lane * 8 + it) ----
    uint4 la = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu), lb = la;
    if(per_point)
    {
      const uint4 *lp = reinterpret_cast<const uint4 *>(lab32 + (size_t)g * 8);
      la = __ldg(lp);
      lb = __ldg(lp + 1);
    }
    const unsigned lw8[8] = { la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w };
    unsigned n = 0;
#pragma unroll
    for(int it = 0; it < 8; it++)
    {
      const unsigned lw = lw8[it];
      const unsigned am4 = (((~lw >> 7) & 0x01010101u) * 0x10204080u) >> 28; // bit j <-> byte j < 128 (a plateau label)
      const unsigned b = __ballot_sync(0xffffffffu, am4 != 0u);
      if(am4)
      {
        const unsigned widx = (unsigned)(lane * 8 + it);
        labs[widx] = lw;
        word_prefetch_l2(FR, wbase + widx);
        act[n + __popc(b & ((1u << lane) - 1u))] = (unsigned short)((widx << 4) | am4);
      }
      n += __popc(b);
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
