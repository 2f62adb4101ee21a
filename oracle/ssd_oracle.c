/*
 * ssd_oracle.c -- TEST INFRASTRUCTURE (see ssd_oracle.h). Plain-C restatement of the reference's per-frame
 * geometry path. Each function cites the reference file:line it follows (paths relative to the upstream
 * repository root). Compile with -O2 -ffp-contract=off (the reference build has no FMA contraction).
 */
#include "ssd_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double x, y; } P2;
typedef struct { int x, y; } P2i;

static int g_sort_ties = 0;

int ssd_oracle_sort_ties(void)
{
  return g_sort_ties;
}

/* ---------------------------------------------------------------------------------------------
 * ProcessingConfiguration / Projection2D  (pointcloud.cpp:60-106, configuration.h:27-52)
 * ------------------------------------------------------------------------------------------- */
int ssd_oracle_derive(const ssd_gpu_config *c, ssd_oracle_derived *d)
{
  d->height_interval_reciprocal = 1.0 / c->height_interval;                                            /* :101 */
  d->min_height = (uint16_t)((c->min_height_above_ground - c->z_min) * d->height_interval_reciprocal); /* :102 */
  d->min_img_y_extent = (int)(c->min_step_depth * c->height / (c->y_max - c->y_min));                  /* :103 */
  d->x_to_image = c->width / (c->x_max - c->x_min);                                                    /* :73 */
  d->y_to_image = c->height / (c->y_max - c->y_min);                                                   /* :74 */
  d->x_to_world = 1 / d->x_to_image;                                                                   /* :75 */
  d->y_to_world = 1 / d->y_to_image;                                                                   /* :76 */
  d->xy_ratio = d->x_to_image / d->y_to_image;                                                         /* :95 */
  d->n_bins = (int)((size_t)((c->z_max - c->z_min) * d->height_interval_reciprocal) + 1);              /* :189,196 */
  d->pad = 0;
  if(d->n_bins < 3 || d->n_bins > SSD_GPU_MAX_BINS)
    return SSD_E_RANGE;
  return SSD_OK;
}

/* ---------------------------------------------------------------------------------------------
 * Transformation_<3>::transform  (transformation.h:59-64): Boost.QVM mat*vec then vec+vec,
 * r_i = ((a_i0*x + a_i1*y) + a_i2*z) + b_i with x,y,z widened from float.
 * ------------------------------------------------------------------------------------------- */
static void camera_to_world(const ssd_gpu_transform *t, const float v[3], double w[3])
{
  const double x = v[0], y = v[1], z = v[2];
  for(int i = 0; i < 3; i++)
    w[i] = ((t->a[i * 3] * x + t->a[i * 3 + 1] * y) + t->a[i * 3 + 2] * z) + t->b[i];
}

int ssd_oracle_camera_to_world(const ssd_gpu_transform *xf, const float *xyz, int n, double *world)
{
  for(int i = 0; i < n; i++)
    camera_to_world(xf, xyz + (size_t)i * 3, world + (size_t)i * 3);
  return SSD_OK;
}

/* ---------------------------------------------------------------------------------------------
 * GeometricTransformation(worldPoints, cameraPoints)  (transformation.cpp:196-215)
 *   _camera = Transformation_<3>(triangleInPlane = cameraPoints)          (:108-157)
 *   _toExternalWorld._world = Transformation_<2>(rp, rpMapping)           (:94-106, makeRotation :65-90)
 * ------------------------------------------------------------------------------------------- */
static void cross3(const double a[3], const double b[3], double r[3])
{
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
}

static void normalized3(const double a[3], double r[3])
{
  const double m2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
  const double rm = 1 / sqrt(m2);
  r[0] = a[0] * rm;
  r[1] = a[1] * rm;
  r[2] = a[2] * rm;
}

static void normalized2(const double a[2], double r[2])
{
  const double m2 = a[0] * a[0] + a[1] * a[1];
  const double rm = 1 / sqrt(m2);
  r[0] = a[0] * rm;
  r[1] = a[1] * rm;
}

int ssd_oracle_make_transform(const double world_pts[9], const double camera_pts[9], ssd_gpu_transform *out)
{
  /* transformation.cpp:113-156 */
  const double *p = camera_pts;
  double u[3], v[3], c[3], n0[3], zb[3], yv[3], yb[3], xb[3];
  for(int i = 0; i < 3; i++)
  {
    u[i] = camera_pts[3 + i] - p[i];
    v[i] = camera_pts[6 + i] - p[i];
  }
  cross3(u, v, c);
  normalized3(c, n0);
  for(int i = 0; i < 3; i++)
    zb[i] = -n0[i];
  yv[0] = 0;
  yv[1] = -zb[2] / zb[1];
  yv[2] = 1;
  normalized3(yv, yb);
  cross3(yb, zb, xb);
  /* _aInv columns = bases; _a = transposed: rows = bases */
  for(int j = 0; j < 3; j++)
  {
    out->a[0 * 3 + j] = xb[j];
    out->a[1 * 3 + j] = yb[j];
    out->a[2 * 3 + j] = zb[j];
  }
  out->b[0] = 0;
  out->b[1] = 0;
  out->b[2] = p[0] * n0[0] + p[1] * n0[1] + p[2] * n0[2];

  /* transformation.cpp:196-212: rp = external world points 0,1 (x,y); rpMapping = cameraToWorld(camera points 0,1) */
  double m0[3], m1[3];
  for(int k = 0; k < 2; k++)
  {
    const double *cp = camera_pts + 3 * k;
    double *m = k ? m1 : m0;
    for(int i = 0; i < 3; i++)
      m[i] = ((out->a[i * 3] * cp[0] + out->a[i * 3 + 1] * cp[1]) + out->a[i * 3 + 2] * cp[2]) + out->b[i];
  }
  /* makeRotation, transformation.cpp:65-90 */
  const double drp[2] = { world_pts[3] - world_pts[0], world_pts[4] - world_pts[1] };
  const double drm[2] = { m1[0] - m0[0], m1[1] - m0[1] };
  double d[2], dm[2];
  normalized2(drp, d);
  normalized2(drm, dm);
  const double xbx = d[0] * dm[0] + d[1] * dm[1];
  const double xby = d[1] * dm[0] - d[0] * dm[1];
  out->ext_a[0] = xbx;
  out->ext_a[1] = -xby;
  out->ext_a[2] = xby;
  out->ext_a[3] = xbx;
  /* _b = rp.front() - _a * rpMapping.front()   (:105) */
  out->ext_b[0] = world_pts[0] - (out->ext_a[0] * m0[0] + out->ext_a[1] * m0[1]);
  out->ext_b[1] = world_pts[1] - (out->ext_a[2] * m0[0] + out->ext_a[3] * m0[1]);
  out->ext_z = world_pts[2]; /* :211 */
  return SSD_OK;
}

/* ---------------------------------------------------------------------------------------------
 * cv::morphologyEx(image, image, MORPH_CLOSE, Mat())  (call sites segmentation.cpp:888,928)
 * OpenCV default: 3x3 rectangle, centre anchor, 1 iteration, border value = identity of the op,
 * i.e. dilate ignores outside pixels, erode ignores outside pixels.
 * ------------------------------------------------------------------------------------------- */
static void box3(const uint8_t *src, uint8_t *dst, int W, int H, int take_max)
{
  for(int y = 0; y < H; y++)
    for(int x = 0; x < W; x++)
    {
      uint8_t m = take_max ? 0 : 255;
      for(int dy = -1; dy <= 1; dy++)
      {
        const int yy = y + dy;
        if(yy < 0 || yy >= H)
          continue;
        for(int dx = -1; dx <= 1; dx++)
        {
          const int xx = x + dx;
          if(xx < 0 || xx >= W)
            continue;
          const uint8_t v = src[(size_t)yy * W + xx];
          if(take_max ? v > m : v < m)
            m = v;
        }
      }
      dst[(size_t)y * W + x] = m;
    }
}

/* binary images only need bit tricks, but keep the general u8 definition; fast path for rows */
static void box3_fast(const uint8_t *src, uint8_t *dst, uint8_t *tmp, int W, int H, int take_max)
{
  if(W < 2 || H < 2)
  {
    box3(src, dst, W, H, take_max);
    return;
  }
#define OP(a, b) (take_max ? ((a) > (b) ? (a) : (b)) : ((a) < (b) ? (a) : (b)))
  for(int y = 0; y < H; y++)
  {
    const uint8_t *s = src + (size_t)y * W;
    uint8_t *t = tmp + (size_t)y * W;
    t[0] = OP(s[0], s[1]);
    for(int x = 1; x < W - 1; x++)
    {
      const uint8_t a = OP(s[x - 1], s[x]);
      t[x] = OP(a, s[x + 1]);
    }
    t[W - 1] = OP(s[W - 2], s[W - 1]);
  }
  for(int y = 0; y < H; y++)
  {
    const uint8_t *t0 = tmp + (size_t)(y > 0 ? y - 1 : y) * W;
    const uint8_t *t1 = tmp + (size_t)y * W;
    const uint8_t *t2 = tmp + (size_t)(y < H - 1 ? y + 1 : y) * W;
    uint8_t *d = dst + (size_t)y * W;
    for(int x = 0; x < W; x++)
    {
      const uint8_t a = OP(t0[x], t1[x]);
      d[x] = OP(a, t2[x]);
    }
  }
#undef OP
}

int ssd_oracle_close(uint8_t *image, int W, int H)
{
  const size_t n = (size_t)W * H;
  uint8_t *dil = (uint8_t *)malloc(n), *tmp = (uint8_t *)malloc(n);
  if(!dil || !tmp)
  {
    free(dil);
    free(tmp);
    return SSD_E_NOMEM;
  }
  box3_fast(image, dil, tmp, W, H, 1);
  box3_fast(dil, image, tmp, W, H, 0);
  free(dil);
  free(tmp);
  return SSD_OK;
}

/* ---------------------------------------------------------------------------------------------
 * Integer lines and the brute-force line fit  (segmentation.cpp:321-487, types.h:117-163)
 * ------------------------------------------------------------------------------------------- */
typedef struct { int a, b, c; } LineI;
typedef struct { double a, b, c; } LineD;

/* LineCoordinates(p,q): a = y2-y1, b = x1-x2, c = x2*y1 - x1*y2   (types.h:143-151) */
static LineI linei_from(P2i p, P2i q)
{
  LineI l = { q.y - p.y, p.x - q.x, q.x * p.y - p.x * q.y };
  return l;
}

static LineD lined_from_pts(P2 p, P2 q)
{
  LineD l = { q.y - p.y, p.x - q.x, q.x * p.y - p.x * q.y };
  return l;
}

static LineD lined_from_i(LineI l)
{
  LineD d = { l.a, l.b, l.c };
  return d;
}

/* ApproximationLine (segmentation.cpp:409-448): residual of the line through points p,q */
static double approximation_residual(LineI l, const P2i *pts, int n, int p, int q, int *dists)
{
  if(n <= 2)
    return 0;
  int m = 0;
  for(int i = 0; i < n; i++)
  {
    if(i == p || i == q)
      continue;
    dists[m++] = abs(pts[i].x * l.a + pts[i].y * l.b + l.c); /* comparableDistance :397-400 */
  }
  int sum = 0;
  const size_t cnt = m > 4 ? (size_t)(m - 1) / 2 : 1; /* :435 */
  for(size_t i = cnt; i > 0; i--)
  {
    int mi = 0;
    for(int j = 1; j < m; j++)
      if(dists[j] < dists[mi])
        mi = j;
    sum += dists[mi];
    dists[mi] = dists[--m]; /* erase (order irrelevant for a multiset minimum) */
  }
  return sum / (cnt * hypot(l.a, l.b)); /* :442 */
}

/* BestLine::findBestLine (segmentation.cpp:461-478): first minimal residual in (p,q) order */
static LineI best_line(const P2i *pts, int n)
{
  int *dists = (int *)malloc(sizeof(int) * (size_t)(n > 2 ? n : 2));
  LineI best = { 0, 0, 0 };
  double bestRes = 0;
  int have = 0;
  for(int p = 0; p < n - 1; p++)
    for(int q = p + 1; q < n; q++)
    {
      const LineI l = linei_from(pts[p], pts[q]);
      const double r = approximation_residual(l, pts, n, p, q, dists);
      if(!have || r < bestRes)
      {
        have = 1;
        bestRes = r;
        best = l;
      }
    }
  free(dists);
  return best;
}

/* FlatLine (segmentation.cpp:490-519) */
typedef struct { double m, n; } FlatLine;
static FlatLine flat_from(LineI l)
{
  FlatLine f = { (double)(-l.a) / l.b, (double)(-l.c) / l.b };
  return f;
}
static P2 flat_point(FlatLine f, double x)
{
  P2 p = { x, x * f.m + f.n };
  return p;
}

/* BoundaryPoints (segmentation.cpp:521-552) */
typedef struct { P2 inner, outer; } Bounds;
static Bounds boundary_points(const P2i *pts, int n, FlatLine line)
{
  Bounds b = { { -1, -1 }, { -1, -1 } };
  const int distanceLimit = 10;
  for(int i = 0; i < n; i++)
  {
    const P2 pd = flat_point(line, pts[i].x);
    if(fabs(pd.y - pts[i].y) < distanceLimit)
    {
      b.inner = pd;
      break;
    }
  }
  for(int i = n - 1; i >= 0; i--)
  {
    const P2 pd = flat_point(line, pts[i].x);
    if(fabs(pd.y - pts[i].y) < distanceLimit)
    {
      b.outer = pd;
      break;
    }
  }
  return b;
}

typedef struct
{
  const P2i *points;
  int n;
  LineI line;
  Bounds bounds;
} HEdge;

static HEdge hedge_make(const P2i *pts, int n)
{
  HEdge e;
  e.points = pts;
  e.n = n;
  e.line = best_line(pts, n);
  e.bounds = boundary_points(pts, n, flat_from(e.line));
  return e;
}

/* ---------------------------------------------------------------------------------------------
 * Scanner (segmentation.cpp:47-157)
 * ------------------------------------------------------------------------------------------- */
typedef struct { int x, yFirst, ySecond; } Scan;

static int probe_vertical(const uint8_t *img, int W, int H, int x, int *yFirst, int *ySecond) /* :85-111 */
{
  for(int yf = 0; yf < H; yf++)
    if(img[(size_t)yf * W + x])
    {
      *yFirst = yf;
      for(int ys = H - 1; ys >= yf; ys--)
        if(img[(size_t)ys * W + x])
        {
          *ySecond = ys;
          return 1;
        }
      return 0;
    }
  return 0;
}

static int scanner_scan(const uint8_t *img, int W, int H, int minImgYExtent, int xStart, int xStep, Scan *out) /* :59-82 */
{
  int n = 0, x = xStart, yf, ys;
  while(probe_vertical(img, W, H, x, &yf, &ys))
  {
    if(ys - yf < minImgYExtent)
      break;
    out[n].x = x;
    out[n].yFirst = yf;
    out[n].ySecond = ys;
    n++;
    x += xStep;
    if(x < 0 || x >= W)
      break;
  }
  return n;
}

typedef struct
{
  P2i *front, *back;
  int n;
} LinePts;

static void lp_push(LinePts *lp, Scan s) /* PointInserter::pushBack :122-126 */
{
  lp->front[lp->n].x = s.x;
  lp->front[lp->n].y = s.ySecond;
  lp->back[lp->n].x = s.x;
  lp->back[lp->n].y = s.yFirst;
  lp->n++;
}

/* Scanner::obtainLinePoints (segmentation.cpp:129-156) */
static void obtain_line_points(const Scan *sl, int nl, const Scan *sr, int nr, LinePts *left, LinePts *right)
{
  const size_t total = (size_t)nl + nr;
  const size_t half = total / 2 + 1;
  size_t indLeft = 0, indRight = 0;
  if((size_t)nl >= half)
  {
    indLeft = nl - half;
    for(int i = (int)indLeft; i >= 0; i--)
      lp_push(right, sl[i]);
  }
  else
  {
    if((size_t)nr > half)
      indRight = nr - half;
    for(int i = (int)indRight; i >= 0; i--)
      lp_push(left, sr[i]);
  }
  for(; indRight < (size_t)nr; indRight++)
    lp_push(right, sr[indRight]);
  for(; indLeft < (size_t)nl; indLeft++)
    lp_push(left, sl[indLeft]);
}

/* ---------------------------------------------------------------------------------------------
 * Line<double> helpers (segmentation.cpp:346-394)
 * ------------------------------------------------------------------------------------------- */
static LineD lined_normalized(LineD l)
{
  const double h = hypot(l.a, l.b);
  LineD r = { l.a / h, l.b / h, l.c / h };
  return r;
}
static LineD linei_normalized(LineI l)
{
  const double h = hypot(l.a, l.b);
  LineD r = { l.a / h, l.b / h, l.c / h };
  return r;
}
static LineD bisector(LineD n, LineD o) /* both already normalized */
{
  LineD r = { n.a + o.a, n.b + o.b, n.c + o.c };
  return r;
}
static LineI linei_reverse(LineI l)
{
  LineI r = { -l.a, -l.b, -l.c };
  return r;
}

/* Line::intersection (segmentation.cpp:350-362): only if the angle exceeds 60 degrees */
static int lined_intersection(LineD t, LineD o, P2 *out)
{
  const double tan60 = 1.7320508075688772935274463415059; /* std::numbers::sqrt3 */
  const double numerator = t.a * o.b - o.a * t.b;          /* det  (types.h:129-132) */
  const double denominator = t.a * o.a + t.b * o.b;
  if(fabs(numerator) > fabs(denominator) * tan60)
  {
    const double det = numerator;
    out->x = (t.b * o.c - o.b * t.c) / det; /* detx (types.h:133-136) */
    out->y = (o.a * t.c - t.a * o.c) / det; /* dety (types.h:137-140) */
    return 1;
  }
  return 0;
}

/* VerticalEdgePointsDetector (segmentation.cpp:243-312) */
static int detect_vertical_points(const uint8_t *img, int W, int x0, int y0, int length, int yEnd, int yStep, int toRight, P2i *out)
{
  int n = 0;
  for(int y = y0; y >= yEnd; y -= yStep)
  {
    const uint8_t *row = img + (size_t)y * W;
    for(int i = length; i > 0; i--)
    {
      const int x = toRight ? x0 + length - i : x0 - length + i;
      if(row[x])
      {
        out[n].x = x;
        out[n].y = y;
        n++;
        break;
      }
    }
  }
  return n;
}

typedef struct { double dist; int idx; } PD;
static int pd_cmp(const void *a, const void *b)
{
  const PD *x = (const PD *)a, *y = (const PD *)b;
  if(x->dist < y->dist)
    return -1;
  if(x->dist > y->dist)
    return 1;
  return x->idx - y->idx;
}

/* VerticalEdgesDetector::findBestPoint (segmentation.cpp:708-728). The reference uses an unstable sort; ties at
 * the selected rank between points that give different edge lines are counted in g_sort_ties. */
static P2i find_best_point(const P2i *pts, int n, LineD base)
{
  PD *d = (PD *)malloc(sizeof(PD) * (size_t)n);
  for(int i = 0; i < n; i++)
  {
    d[i].dist = fabs(pts[i].x * base.a + pts[i].y * base.b + base.c);
    d[i].idx = i;
  }
  qsort(d, (size_t)n, sizeof(PD), pd_cmp);
  const size_t best = 2 * (size_t)n / 3;
  const P2i r = pts[d[best].idx];
  for(int i = 0; i < n; i++)
    if(i != (int)best && d[i].dist == d[best].dist)
    {
      const P2i o = pts[d[i].idx];
      if(-base.a * o.x - base.b * o.y != -base.a * r.x - base.b * r.y)
      {
        g_sort_ties++;
        break;
      }
    }
  free(d);
  return r;
}

/* VerticalEdgesDetector::detectEdge (segmentation.cpp:681-706) */
static int detect_vertical_edge(const uint8_t *img, int W, int H, P2 front, P2 back, LineD base, int xExtension, int isLeft, LineD *edge)
{
  const int yStep = 10;
  int left = (int)((front.x < back.x ? front.x : back.x) - xExtension);  /* std::min(front.x, back.x) */
  int right = (int)((front.x < back.x ? back.x : front.x) + xExtension); /* std::max(front.x, back.x) */
  int yStart = (int)(front.y - yStep);
  int yEnd = (int)(back.y + yStep);
  if(left < 0)
    left = 0;
  if(right >= W)
    right = W - 1;
  if(yStart >= H)
    yStart = H - 1;
  if(yEnd < 0)
    yEnd = 0;
  if(yStart < yEnd)
    return 0;
  const int cap = (yStart - yEnd) / yStep + 1;
  P2i *pts = (P2i *)malloc(sizeof(P2i) * (size_t)cap);
  const int n = isLeft ? detect_vertical_points(img, W, left, yStart, right - left, yEnd, yStep, 1, pts)
                       : detect_vertical_points(img, W, right, yStart, right - left, yEnd, yStep, 0, pts);
  if(n == 0)
  {
    free(pts);
    return 0;
  }
  const P2i bp = find_best_point(pts, n, base);
  free(pts);
  /* Line::parallel (:367-371) */
  edge->a = base.a;
  edge->b = base.b;
  edge->c = -base.a * bp.x - base.b * bp.y;
  return 1;
}

/* Quadrilateral::isConvex (segmentation.cpp:755-787) */
static int is_convex(const P2 q[4])
{
  const P2 v[4] = { { q[1].x - q[0].x, q[1].y - q[0].y },
                    { q[3].x - q[1].x, q[3].y - q[1].y },
                    { q[2].x - q[3].x, q[2].y - q[3].y },
                    { q[0].x - q[2].x, q[0].y - q[2].y } };
#define POS(i, j) (v[i].x * v[j].y - v[j].x * v[i].y > 0)
  const int positive = POS(0, 1);
  return positive == POS(1, 2) && positive == POS(2, 3) && positive == POS(3, 0);
#undef POS
}

/* Segmentation::detectOutline after the close (segmentation.cpp:930-946) */
static void detect_outline_closed(const uint8_t *img, int W, int H, int minImgYExtent, double xyRatio, P2 quad[4], int *valid)
{
  for(int i = 0; i < 4; i++)
    quad[i].x = quad[i].y = 0; /* Outline{} value-initialises the quadrilateral */
  *valid = 0;
  const int xStep = 25; /* HorizontalEdgesDetector::xStep :605 */
  const int xCenter = W / 2;
  const int cap = W / xStep + 2;
  Scan *sr = (Scan *)malloc(sizeof(Scan) * (size_t)cap), *sl = (Scan *)malloc(sizeof(Scan) * (size_t)cap);
  /* HorizontalEdgesDetector::detect :607-620 */
  const int nr = scanner_scan(img, W, H, minImgYExtent, xCenter, xStep, sr);
  int nl = 0;
  int ok = 0;
  if(nr > 0)
  {
    nl = scanner_scan(img, W, H, minImgYExtent, xCenter - xStep, -xStep, sl);
    ok = nl + nr >= 3;
  }
  if(ok)
  {
    const int total = nl + nr;
    P2i *buf = (P2i *)malloc(sizeof(P2i) * (size_t)(4 * (total + 2)));
    LinePts left = { buf, buf + (total + 2), 0 }, right = { buf + 2 * (total + 2), buf + 3 * (total + 2), 0 };
    obtain_line_points(sl, nl, sr, nr, &left, &right);
    /* HorizontalEdges :592-599 */
    const HEdge frontLeft = hedge_make(left.front, left.n);
    const HEdge frontRight = hedge_make(right.front, right.n);
    const HEdge backLeft = hedge_make(left.back, left.n);
    const HEdge backRight = hedge_make(right.back, right.n);

    /* VerticalEdgesDetector::calcBaseLine :672-679 */
    const LineD front = bisector(linei_normalized(linei_reverse(frontLeft.line)), linei_normalized(frontRight.line));
    const LineD back = bisector(linei_normalized(linei_reverse(backLeft.line)), linei_normalized(backRight.line));
    const LineD center = bisector(lined_normalized(front), lined_normalized(back));
    const double cf = xyRatio * xyRatio;
    const LineD corr = { center.a * cf, center.b, center.c }; /* slopeCorrection :377-380 */
    const P2i p0 = frontLeft.points[0];
    const LineD base = { -corr.b, corr.a, corr.b * p0.x - corr.a * p0.y }; /* perpendicular :372-376 */

    LineD le, re;
    const int hasLeft = detect_vertical_edge(img, W, H, frontLeft.bounds.outer, backLeft.bounds.outer, base, xStep, 1, &le);
    const int hasRight = detect_vertical_edge(img, W, H, frontRight.bounds.outer, backRight.bounds.outer, base, xStep, 0, &re);

    /* Corners :738-750 and value_or fall-backs :939-945 */
    quad[0] = frontLeft.bounds.outer;
    quad[1] = frontRight.bounds.outer;
    quad[2] = backLeft.bounds.outer;
    quad[3] = backRight.bounds.outer;
    P2 c;
    if(hasLeft)
    {
      if(lined_intersection(le, lined_from_i(frontLeft.line), &c))
        quad[0] = c;
      if(lined_intersection(le, lined_from_i(backLeft.line), &c))
        quad[2] = c;
    }
    if(hasRight)
    {
      if(lined_intersection(re, lined_from_i(frontRight.line), &c))
        quad[1] = c;
      if(lined_intersection(re, lined_from_i(backRight.line), &c))
        quad[3] = c;
    }
    *valid = is_convex(quad);
    free(buf);
  }
  free(sr);
  free(sl);
}

int ssd_oracle_detect_outline(const uint8_t *image, int W, int H, int minImgYExtent, double xyRatio, double quad_px[8], int *valid)
{
  uint8_t *img = (uint8_t *)malloc((size_t)W * H);
  if(!img)
    return SSD_E_NOMEM;
  memcpy(img, image, (size_t)W * H);
  ssd_oracle_close(img, W, H); /* segmentation.cpp:928 */
  P2 q[4];
  detect_outline_closed(img, W, H, minImgYExtent, xyRatio, q, valid);
  for(int i = 0; i < 4; i++)
  {
    quad_px[i * 2] = q[i].x;
    quad_px[i * 2 + 1] = q[i].y;
  }
  free(img);
  return SSD_OK;
}

/* ---------------------------------------------------------------------------------------------
 * BottomScanner + Segmentation::detectFrontEdge (segmentation.cpp:159-241, 879-917)
 * ------------------------------------------------------------------------------------------- */
static int probe_bottom_up(const uint8_t *img, int W, int H, int x, int *yEdge) /* :225-240 */
{
  const int yStop = H / 2;
  for(int y = H - 1; y > yStop; y--)
    if(img[(size_t)y * W + x])
    {
      *yEdge = y;
      return 1;
    }
  return 0;
}

static int bottom_scan(const uint8_t *img, int W, int H, int xStart, int xStep, P2i *pts) /* :169-222 */
{
  int n = 0, x = xStart, y;
  do
  {
    if(probe_bottom_up(img, W, H, x, &y))
    {
      pts[n].x = x;
      pts[n].y = y;
      n++;
      break;
    }
    x += xStep;
  } while(x < W);
  if(n == 0)
  {
    x = xStart - xStep;
    do
    {
      if(probe_bottom_up(img, W, H, x, &y))
      {
        pts[n].x = x;
        pts[n].y = y;
        n++;
        break;
      }
      x -= xStep;
    } while(x >= 0);
    if(n == 0)
      return 0;
  }
  xStart = x;
  x += xStep;
  while(x < W && probe_bottom_up(img, W, H, x, &y))
  {
    pts[n].x = x;
    pts[n].y = y;
    n++;
    x += xStep;
  }
  x = xStart - xStep;
  while(x >= 0 && probe_bottom_up(img, W, H, x, &y))
  {
    pts[n].x = x;
    pts[n].y = y;
    n++;
    x -= xStep;
  }
  return n;
}

static void detect_front_edge_closed(const uint8_t *img, int W, int H, P2 *left, P2 *right, int *valid)
{
  left->x = left->y = right->x = right->y = 0;
  *valid = 0;
  const int xStep = 50;
  const int xCenter = W / 2;
  P2i *pts = (P2i *)malloc(sizeof(P2i) * (size_t)(W / xStep + 3));
  const int n = bottom_scan(img, W, H, xCenter, xStep, pts);
  if(n >= 2)
  {
    const FlatLine edge = flat_from(best_line(pts, n));
    /* ranges::minmax by x (:897-900): first smallest, last largest */
    int lo = 0, hi = 0;
    for(int i = 1; i < n; i++)
    {
      if(pts[i].x < pts[lo].x)
        lo = i;
      if(!(pts[i].x < pts[hi].x))
        hi = i;
    }
    *left = flat_point(edge, pts[lo].x);
    *right = flat_point(edge, pts[hi].x);
    *valid = 1;
  }
  free(pts);
}

int ssd_oracle_detect_front_edge(const uint8_t *image, int W, int H, double left_px[2], double right_px[2], int *valid)
{
  uint8_t *img = (uint8_t *)malloc((size_t)W * H);
  if(!img)
    return SSD_E_NOMEM;
  memcpy(img, image, (size_t)W * H);
  ssd_oracle_close(img, W, H); /* segmentation.cpp:888 */
  P2 l, r;
  detect_front_edge_closed(img, W, H, &l, &r, valid);
  left_px[0] = l.x;
  left_px[1] = l.y;
  right_px[0] = r.x;
  right_px[1] = r.y;
  free(img);
  return SSD_OK;
}

/* ---------------------------------------------------------------------------------------------
 * QuadrilateralTest (quadrilateralTest.cpp:29-118 Sector/BBox, :123-259 segments, :275-451 ctor + test)
 * ------------------------------------------------------------------------------------------- */
typedef struct { double lower, upper; } Sector;
static Sector sector_make(double a, double b) /* :29-41 */
{
  Sector s = { a, a };
  if(s.lower > b)
    s.lower = b;
  else if(s.upper < b)
    s.upper = b;
  return s;
}
static void sector_expand(Sector *s, double c)
{
  if(s->lower > c)
    s->lower = c;
  else if(s->upper < c)
    s->upper = c;
}
static int sector_overlaps(Sector a, Sector b) { return a.lower < b.upper && a.upper > b.lower; }
static int sector_is_above(Sector a, Sector b) { return (a.lower + a.upper) / 2 < b.lower; }
static int sector_is_below(Sector a, Sector b) { return (a.lower + a.upper) / 2 > b.upper; }
static int sector_within(Sector a, double c) { return a.lower < c && c < a.upper; }

typedef struct { Sector x, y; } BBox;
enum { REL_NOWHERE, REL_XABOVE, REL_XBELOW, REL_YABOVE, REL_YBELOW, REL_MAX };
static int bbox_relpos(BBox a, BBox o) /* :96-113 */
{
  if(sector_overlaps(a.y, o.y))
  {
    if(sector_is_above(a.x, o.x))
      return REL_XABOVE;
    if(sector_is_below(a.x, o.x))
      return REL_XBELOW;
  }
  if(sector_overlaps(a.x, o.x))
  {
    if(sector_is_above(a.y, o.y))
      return REL_YABOVE;
    if(sector_is_below(a.y, o.y))
      return REL_YBELOW;
  }
  return REL_NOWHERE;
}

typedef struct
{
  BBox box;
  int steep;       /* SteepLine vs FlatLine */
  int left_is_pos; /* isLeft(p) == isPositive(p) ? */
  double k, c;     /* FlatLine: x*k + y + c ; SteepLine: x + y*k + c */
} Segment;

static Segment segment_create(P2 p, P2 q) /* :240-259, lines :133-167 */
{
  Segment s;
  s.box.x = sector_make(p.x, q.x);
  s.box.y = sector_make(p.y, q.y);
  const double dx = q.x - p.x, dy = q.y - p.y;
  const LineD l = lined_from_pts(p, q);
  if(fabs(dx) < fabs(dy))
  {
    s.steep = 1;
    s.k = l.b / l.a;
    s.c = l.c / l.a;
    s.left_is_pos = !(dy > 0); /* SteepPositive: !isPositive ; SteepNegative: isPositive */
  }
  else
  {
    s.steep = 0;
    s.k = l.a / l.b;
    s.c = l.c / l.b;
    s.left_is_pos = dx > 0; /* FlatPositive: isPositive ; FlatNegative: !isPositive */
  }
  return s;
}

static int segment_is_left(const Segment *s, P2 p)
{
  const int pos = s->steep ? (p.x + p.y * s->k + s->c > 0) : (p.x * s->k + p.y + s->c > 0);
  return s->left_is_pos ? pos : !pos;
}

typedef struct
{
  double upperX;
  int nseg;
  int seg[4];
  int neighbor[REL_MAX];
} Cell;
typedef struct
{
  double upperY;
  int ncell;
  Cell cell[3];
} Row;
typedef struct
{
  BBox total;
  Segment seg[4];
  int inside_is_left;
  int nrow;
  Row row[3];
} QuadTest;

static int cell_same_segs(const Cell *a, const Cell *b)
{
  if(a->nseg != b->nseg)
    return 0;
  for(int i = 0; i < a->nseg; i++)
    if(a->seg[i] != b->seg[i])
      return 0;
  return 1;
}

/* returns 0 ok, 1 if the reference ctor would throw std::invalid_argument */
static int quadtest_init(QuadTest *t, const P2 q[4])
{
  /* _totalBox(q): BBox(q0,q1) expanded by q2,q3 (:78-83) */
  t->total.x = sector_make(q[0].x, q[1].x);
  t->total.y = sector_make(q[0].y, q[1].y);
  sector_expand(&t->total.x, q[2].x);
  sector_expand(&t->total.y, q[2].y);
  sector_expand(&t->total.x, q[3].x);
  sector_expand(&t->total.y, q[3].y);
  t->seg[0] = segment_create(q[0], q[1]);
  t->seg[1] = segment_create(q[1], q[3]);
  t->seg[2] = segment_create(q[3], q[2]);
  t->seg[3] = segment_create(q[2], q[0]);
  t->inside_is_left = segment_is_left(&t->seg[0], q[3]);
  if(t->inside_is_left != segment_is_left(&t->seg[1], q[2]) || t->inside_is_left != segment_is_left(&t->seg[2], q[0]) ||
     t->inside_is_left != segment_is_left(&t->seg[3], q[1]))
    return 1; /* :283-288 */

  double xs[4] = { q[0].x, q[1].x, q[2].x, q[3].x }, ys[4] = { q[0].y, q[1].y, q[2].y, q[3].y };
  for(int i = 1; i < 4; i++) /* ranges::sort of four doubles */
    for(int j = i; j > 0 && xs[j] < xs[j - 1]; j--)
    {
      const double tmp = xs[j];
      xs[j] = xs[j - 1];
      xs[j - 1] = tmp;
    }
  for(int i = 1; i < 4; i++)
    for(int j = i; j > 0 && ys[j] < ys[j - 1]; j--)
    {
      const double tmp = ys[j];
      ys[j] = ys[j - 1];
      ys[j - 1] = tmp;
    }

  t->nrow = 0;
  double lowerY = ys[0];
  for(int yi = 1; yi < 4; yi++) /* :314-345 */
  {
    if(lowerY < ys[yi])
    {
      Row *row = &t->row[t->nrow++];
      row->upperY = ys[yi];
      row->ncell = 0;
      double lowerX = xs[0];
      for(int xi = 1; xi < 4; xi++)
      {
        if(lowerX < xs[xi])
        {
          Cell *cell = &row->cell[row->ncell++];
          memset(cell, 0, sizeof(*cell));
          cell->upperX = xs[xi];
          BBox cb;
          cb.x = sector_make(lowerX, cell->upperX);
          cb.y = sector_make(lowerY, row->upperY);
          for(int si = 0; si < 4; si++)
          {
            const BBox sb = t->seg[si].box;
            if(sector_overlaps(cb.x, sb.x) && sector_overlaps(cb.y, sb.y))
              cell->seg[cell->nseg++] = si;
            if(cell->nseg == 0)
              cell->neighbor[bbox_relpos(cb, sb)] = 1;
          }
          lowerX = cell->upperX;
        }
      }
      lowerY = row->upperY;
    }
  }
  if(t->nrow == 0)
    return 1; /* :347-348 */
  for(int r = 0; r < t->nrow; r++)
  {
    if(t->row[r].ncell == 0)
      return 1; /* :352-353 */
    for(int c = 0; c < t->row[r].ncell; c++)
      if(t->row[r].cell[c].nseg > 2)
        return 1; /* :357-358 */
  }
  for(int r = 0; r < t->nrow; r++) /* :362-380 */
  {
    Row *row = &t->row[r];
    int ci = 0;
    while(ci != row->ncell - 1)
    {
      const Cell *cur = &row->cell[ci], *nxt = &row->cell[ci + 1];
      if(cur->nseg == 0 && nxt->nseg == 0)
        return 1;
      if(cur->nseg > 1 && nxt->nseg > 1)
        return 1;
      if(cell_same_segs(cur, nxt))
      {
        for(int k = ci; k < row->ncell - 1; k++)
          row->cell[k] = row->cell[k + 1];
        row->ncell--;
      }
      else
        ci++;
    }
  }
  return 0;
}

static int quadtest_within(const QuadTest *t, P2 p) /* :445-451, selectors :494-598, testers :453-492 */
{
  if(!(sector_within(t->total.x, p.x) && sector_within(t->total.y, p.y)))
    return 0;
  const Row *row = &t->row[0];
  if(t->nrow == 2)
    row = p.y < t->row[0].upperY ? &t->row[0] : &t->row[1];
  else if(t->nrow == 3)
    row = p.y < t->row[0].upperY ? &t->row[0] : (p.y < t->row[1].upperY ? &t->row[1] : &t->row[2]);
  const Cell *cell = &row->cell[0];
  if(row->ncell == 2)
    cell = p.x < row->cell[0].upperX ? &row->cell[0] : &row->cell[1];
  else if(row->ncell == 3)
    cell = p.x < row->cell[0].upperX ? &row->cell[0] : (p.x < row->cell[1].upperX ? &row->cell[1] : &row->cell[2]);
  switch(cell->nseg)
  {
    case 0:
      return cell->neighbor[REL_XABOVE] && cell->neighbor[REL_XBELOW] && cell->neighbor[REL_YABOVE] && cell->neighbor[REL_YBELOW];
    case 1:
      return segment_is_left(&t->seg[cell->seg[0]], p) == t->inside_is_left;
    default:
      return segment_is_left(&t->seg[cell->seg[0]], p) == t->inside_is_left &&
             segment_is_left(&t->seg[cell->seg[1]], p) == t->inside_is_left;
  }
}

int ssd_oracle_points_in_quad(const double quad[8], const double *xy, int n, uint8_t *inside, int *ctor_status)
{
  P2 q[4];
  for(int i = 0; i < 4; i++)
  {
    q[i].x = quad[i * 2];
    q[i].y = quad[i * 2 + 1];
  }
  memset(inside, 0, (size_t)n);
  QuadTest t;
  *ctor_status = quadtest_init(&t, q);
  if(*ctor_status)
    return SSD_OK;
  for(int i = 0; i < n; i++)
  {
    const P2 p = { xy[i * 2], xy[i * 2 + 1] };
    inside[i] = (uint8_t)quadtest_within(&t, p);
  }
  return SSD_OK;
}

/* ---------------------------------------------------------------------------------------------
 * Vertical faces from the remainder. The reference stops at "TODO use remainder to detect vertical faces"
 * (pointcloud.cpp:293) after collecting the remainder (:283-291): there is no reference behaviour to restate, so this is the
 * restatement of the DEFINITION in include/ssd_gpu.h (ssd_gpu_riser) -- parity for this row is unpinned by construction.
 * labels / plateaus: as ssd_oracle_process (or the compiled reference, whose remainder vector holds exactly the points with
 * label SSD_LABEL_REMAINDER) produced them for the same vertices.
 * ------------------------------------------------------------------------------------------- */
int ssd_oracle_vertical_faces(const ssd_gpu_config *cfg, const ssd_gpu_transform *xf, const float *xyz, const uint8_t *labels,
                              const ssd_gpu_plateau *plateaus, int n_plateaus, ssd_gpu_riser *out, int cap, int *n)
{
  ssd_oracle_derived d;
  const int rc = ssd_oracle_derive(cfg, &d);
  if(rc)
    return rc;
  const int R = n_plateaus > 1 ? n_plateaus - 1 : 0;
  if(n)
    *n = R;
  if(R > SSD_GPU_MAX_PLATEAUS)
    return SSD_E_RANGE;
  uint32_t cnt[SSD_GPU_MAX_PLATEAUS];
  long long sx[SSD_GPU_MAX_PLATEAUS], sy[SSD_GPU_MAX_PLATEAUS], xmin[SSD_GPU_MAX_PLATEAUS], xmax[SSD_GPU_MAX_PLATEAUS], ymin[SSD_GPU_MAX_PLATEAUS],
    ymax[SSD_GPU_MAX_PLATEAUS];
  for(int k = 0; k < SSD_GPU_MAX_PLATEAUS; k++)
  {
    cnt[k] = 0;
    sx[k] = sy[k] = 0;
    xmin[k] = ymin[k] = 0x7fffffff;
    xmax[k] = ymax[k] = -1;
  }
  const size_t N = (size_t)cfg->width * cfg->height;
  for(size_t i = 0; i < N; i++)
  {
    if(labels[i] != SSD_LABEL_REMAINDER)
      continue;
    double w[3];
    camera_to_world(xf, xyz + i * 3, w);
    const int h = (uint16_t)((w[2] - cfg->z_min) * d.height_interval_reciprocal); /* calcHeights (pointcloud.cpp:175) */
    int r = -1;
    for(int k = 0; k + 1 < n_plateaus; k++)
      if(h > plateaus[k].hmax && h < plateaus[k + 1].hmin)
        r = k;
    if(r < 0)
      continue;
    const long long X = (long long)((w[0] - cfg->x_min) * 65536.0), Y = (long long)((w[1] - cfg->y_min) * 65536.0);
    cnt[r]++;
    sx[r] += X;
    sy[r] += Y;
    if(X < xmin[r]) xmin[r] = X;
    if(X > xmax[r]) xmax[r] = X;
    if(Y < ymin[r]) ymin[r] = Y;
    if(Y > ymax[r]) ymax[r] = Y;
  }
  for(int k = 0; k < R && k < cap && out; k++)
  {
    ssd_gpu_riser o;
    memset(&o, 0, sizeof o);
    o.lower_plateau = k;
    o.upper_plateau = k + 1;
    o.n_points = cnt[k];
    o.z_bottom = cfg->z_min + (double)(plateaus[k].hmax + 1) * cfg->height_interval;
    o.z_top = cfg->z_min + (double)plateaus[k + 1].hmin * cfg->height_interval;
    if(cnt[k])
    {
      o.x_min = cfg->x_min + (double)xmin[k] / 65536.0;
      o.x_max = cfg->x_min + (double)xmax[k] / 65536.0;
      o.y_min = cfg->y_min + (double)ymin[k] / 65536.0;
      o.y_max = cfg->y_min + (double)ymax[k] / 65536.0;
      o.x_mean = cfg->x_min + (double)(unsigned long long)sx[k] / (double)cnt[k] / 65536.0;
      o.y_mean = cfg->y_min + (double)(unsigned long long)sy[k] / (double)cnt[k] / 65536.0;
    }
    out[k] = o;
  }
  return SSD_OK;
}

/* ---------------------------------------------------------------------------------------------
 * The frame pipeline (pointcloud.cpp:108-626)
 * ------------------------------------------------------------------------------------------- */
typedef struct
{
  const ssd_gpu_config *cfg;
  ssd_oracle_derived d;
  /* in-range points (pointcloud.cpp:150-178) */
  uint32_t n_points;
  double *world;   /* n_points x 3 */
  uint32_t *pixel; /* pixel index of each in-range point */
  uint16_t *height;
} Frame;

/* StairsDetector::projectToBinaryImage (pointcloud.cpp:458-471) with Projection2D::worldToImage (:79-83) */
static int project_to_binary_image(const Frame *f, const uint32_t *idx, uint32_t n, uint8_t *img)
{
  const int W = f->cfg->width, H = f->cfg->height;
  int oob = 0;
  memset(img, 0, (size_t)W * H);
  for(uint32_t i = 0; i < n; i++)
  {
    const double *p = f->world + (size_t)idx[i] * 3;
    const int ix = (int)((p[0] - f->cfg->x_min) * f->d.x_to_image);
    const int iy = (int)((f->cfg->y_max - p[1]) * f->d.y_to_image);
    const long long off = (long long)iy * W + ix; /* Mat::ptr(y, x) has no bounds check */
    if(off >= 0 && off < (long long)W * H)
      img[off] = 0xff;
    else
      oob = 1;
  }
  return oob;
}

static P2 image_to_world(const Frame *f, P2 p) /* Projection2D::imageToWorld (:84-88) */
{
  P2 w = { f->cfg->x_min + p.x * f->d.x_to_world, f->cfg->y_max - p.y * f->d.y_to_world };
  return w;
}

/* getPointsInQuadrilateral + calcAverageZ (pointcloud.cpp:560-581); returns ctor status */
static int points_in_quad_mean(const Frame *f, const uint32_t *idx, uint32_t n, const P2 q[4], uint32_t *sel, uint32_t *n_sel,
                               double *mean)
{
  QuadTest t;
  *n_sel = 0;
  *mean = 0;
  if(quadtest_init(&t, q))
    return 1;
  double sum = 0;
  uint32_t m = 0;
  for(uint32_t i = 0; i < n; i++)
  {
    const double *p = f->world + (size_t)idx[i] * 3;
    const P2 p2 = { p[0], p[1] };
    if(quadtest_within(&t, p2))
    {
      sum = sum + p[2];
      if(sel)
        sel[m] = idx[i];
      m++;
    }
  }
  *n_sel = m;
  *mean = sum / m; /* NaN for an empty set, as the reference (:580) */
  return 0;
}

/* StairsDetector::calcGroundQuadrilateral (pointcloud.cpp:489-512) */
static void calc_ground_quadrilateral(const Frame *f, const P2 q[4], P2 gq[4])
{
  const double yMin = f->cfg->y_min;
#define CALCDX(p, q) (((q).y - yMin) * ((q).y - (p).y) / ((q).x - (p).x))
  if(q[0].y < q[1].y)
  {
    gq[0].x = q[0].x;
    gq[0].y = yMin;
    gq[1].x = q[1].x + CALCDX(q[0], q[1]);
    gq[1].y = yMin;
  }
  else
  {
    gq[0].x = q[0].x + CALCDX(q[1], q[0]);
    gq[0].y = yMin;
    gq[1].x = q[1].x;
    gq[1].y = yMin;
  }
#undef CALCDX
  gq[2] = q[0];
  gq[3] = q[1];
}

/* nested Line::intersection without the angle gate (pointcloud.cpp:514-526) */
static P2 line_intersection_plain(LineD t, LineD o)
{
  const double d = t.a * o.b - o.a * t.b;
  P2 r = { (t.b * o.c - o.b * t.c) / d, (o.a * t.c - t.a * o.c) / d };
  return r;
}

/* world quadrilaterals {x, y, averageZ} of the steps of the last ssd_oracle_process on this thread, before
 * ToExternalWorld: what detectStairs hands to drawStairStep (pointcloud.cpp:367-368) */
static __thread double g_last_step_quads[SSD_GPU_MAX_STEPS][4][3];
static __thread int g_last_n_steps;

int ssd_oracle_process(const ssd_gpu_config *cfg, const ssd_gpu_transform *xf, const float *xyz, uint8_t *labels, uint32_t *hist_out,
                       int hist_cap, ssd_gpu_frame_info *info, ssd_gpu_plateau *plats, ssd_gpu_step *steps)
{
  Frame f;
  memset(&f, 0, sizeof(f));
  f.cfg = cfg;
  int rc = ssd_oracle_derive(cfg, &f.d);
  if(rc)
    return rc;
  const int W = cfg->width, H = cfg->height;
  const size_t N = (size_t)W * H;
  memset(info, 0, sizeof(*info));
  info->ground_index = info->first_valid_index = -1;
  info->n_bins = f.d.n_bins;

  f.world = (double *)malloc(sizeof(double) * 3 * N);
  f.pixel = (uint32_t *)malloc(sizeof(uint32_t) * N);
  f.height = (uint16_t *)malloc(sizeof(uint16_t) * N);
  uint32_t *cur = (uint32_t *)malloc(sizeof(uint32_t) * N);   /* pointsHt (indices into f.world) */
  uint32_t *upper = (uint32_t *)malloc(sizeof(uint32_t) * N);
  uint32_t *store = (uint32_t *)malloc(sizeof(uint32_t) * N); /* plateau point lists, concatenated */
  uint32_t *sel = (uint32_t *)malloc(sizeof(uint32_t) * N);
  uint8_t *img = (uint8_t *)malloc(N);
  if(!f.world || !f.pixel || !f.height || !cur || !upper || !store || !sel || !img)
  {
    rc = SSD_E_NOMEM;
    goto done;
  }

  /* PointsExtraction::extract (pointcloud.cpp:122-178): z>0, transform, range, height index */
  uint32_t n = 0, nNonZero = 0;
  for(size_t i = 0; i < N; i++)
  {
    const float *v = xyz + i * 3;
    if(!(v[2] > 0))
    {
      if(labels)
        labels[i] = SSD_LABEL_INVALID;
      continue;
    }
    nNonZero++;
    double w[3];
    camera_to_world(xf, v, w);
    const int in = w[0] > cfg->x_min && w[0] < cfg->x_max && w[1] > cfg->y_min && w[1] < cfg->y_max && w[2] > cfg->z_min &&
                   w[2] < cfg->z_max;
    if(labels)
      labels[i] = in ? SSD_LABEL_REMAINDER : SSD_LABEL_OUT_OF_RANGE;
    if(!in)
      continue;
    f.world[(size_t)n * 3] = w[0];
    f.world[(size_t)n * 3 + 1] = w[1];
    f.world[(size_t)n * 3 + 2] = w[2];
    f.pixel[n] = (uint32_t)i;
    f.height[n] = (uint16_t)((w[2] - cfg->z_min) * f.d.height_interval_reciprocal); /* :175 */
    n++;
  }
  f.n_points = n;
  info->n_nonzero = nNonZero;
  info->n_in_range = n;

  /* HeightsHistogram::calcHist (:194-204) */
  uint32_t hist[SSD_GPU_MAX_BINS + 2];
  memset(hist, 0, sizeof(hist));
  for(uint32_t i = 0; i < n; i++)
    ++hist[f.height[i]];
  if(hist_out)
    for(int i = 0; i < hist_cap && i < f.d.n_bins; i++)
      hist_out[i] = hist[i];

  /* findPeaks (:214-241) + filterPeaks (:243-256) */
  uint16_t peaks[SSD_GPU_MAX_BINS];
  int nPeaks = 0;
  {
    const size_t last = (size_t)f.d.n_bins - 1;
    int ascending = 0;
    for(uint16_t i = 0; i < last; i++)
    {
      const uint32_t c = hist[i], s = hist[i + 1];
      if(c < s)
      {
        ascending = 1;
        continue;
      }
      if(c > s)
      {
        if(ascending)
        {
          const uint32_t np = hist[i];
          if(!(np < cfg->min_peak_points) && (uint32_t)((np * 2 - hist[i - 1] - hist[i + 1]) * 2) > np)
            peaks[nPeaks++] = i;
        }
        ascending = 0;
      }
    }
  }
  if(nPeaks > SSD_GPU_MAX_PLATEAUS)
  {
    info->status |= SSD_STATUS_TOO_MANY_PLATEAUS;
    nPeaks = SSD_GPU_MAX_PLATEAUS;
  }
  info->n_plateaus = nPeaks;

  /* PlateausExtraction::extractPlateaus (:280-343), literally: two stable splits per peak */
  uint32_t nCur = n;
  for(uint32_t i = 0; i < n; i++)
    cur[i] = i;
  uint32_t *pStart[SSD_GPU_MAX_PLATEAUS];
  uint32_t pCount[SSD_GPU_MAX_PLATEAUS];
  uint32_t stored = 0;
  for(int k = 0; k < nPeaks; k++)
  {
    const uint16_t h = peaks[k];
    const uint16_t pred = (uint16_t)(h - 1), succ = (uint16_t)(h + 1);
    uint16_t hmin, hmax;
    if(hist[pred] > hist[succ])
    {
      hmin = pred;
      hmax = h;
    }
    else
    {
      hmin = h;
      hmax = succ;
    }
    const uint16_t thr = (uint16_t)(hmin - 1); /* wraps to 65535 when hmin == 0 (:324,337) */
    if(hmin == 0)
      info->status |= SSD_STATUS_HMIN_WRAP;
    uint32_t nUpper = 0;
    for(uint32_t i = 0; i < nCur; i++)
      if(thr < f.height[cur[i]])
        upper[nUpper++] = cur[i];
    pStart[k] = store + stored;
    uint32_t np = 0;
    nCur = 0;
    for(uint32_t i = 0; i < nUpper; i++)
    {
      if(hmax < f.height[upper[i]])
        cur[nCur++] = upper[i];
      else
        pStart[k][np++] = upper[i];
    }
    pCount[k] = np;
    stored += np;
    ssd_gpu_plateau *o = &plats[k];
    memset(o, 0, sizeof(*o));
    o->height = h;
    o->hmin = hmin;
    o->hmax = hmax;
    o->n_points = np;
    o->quad_status = -1;
    if(labels)
      for(uint32_t i = 0; i < np; i++)
        labels[f.pixel[pStart[k][i]]] = (uint8_t)k;
  }

  /* StairsDetector::detectStairSteps (:399-456) */
  P2 quadW[SSD_GPU_MAX_PLATEAUS][4];
  memset(quadW, 0, sizeof(quadW));
  int groundInd = -1, firstValid = -1;
  {
    size_t maxGround = 0;
    int i = 0;
    for(; i < nPeaks; i++)
    {
      if(plats[i].height >= f.d.min_height)
        break;
      if(maxGround < pCount[i])
      {
        maxGround = pCount[i];
        groundInd = i;
      }
    }
    for(; i < nPeaks; i++)
    {
      if(project_to_binary_image(&f, pStart[i], pCount[i], img))
        info->status |= SSD_STATUS_BEV_OOB;
      ssd_oracle_close(img, W, H);
      P2 qpx[4];
      int valid;
      detect_outline_closed(img, W, H, f.d.min_img_y_extent, f.d.xy_ratio, qpx, &valid);
      for(int c = 0; c < 4; c++)
        quadW[i][c] = image_to_world(&f, qpx[c]); /* imgPointsToWorld :476-487 */
      plats[i].valid = valid;
      plats[i].outlined = 1;
      if(valid && firstValid < 0)
        firstValid = i;
    }
  }
  info->ground_index = groundInd;
  info->first_valid_index = firstValid;

  typedef struct { double x, y, z; } P3;
  P3 stepQuads[SSD_GPU_MAX_STEPS][4];
  int nSteps = 0;
  if(firstValid >= 0)
  {
    if(groundInd >= 0)
    {
      /* ground (:436-443) + calcGround (:528-547) */
      calc_ground_quadrilateral(&f, quadW[firstValid], quadW[groundInd]);
      plats[groundInd].valid = 1;
      const P2 *gq = quadW[groundInd];
      uint32_t nSel;
      double mean;
      const int st = points_in_quad_mean(&f, pStart[groundInd], pCount[groundInd], gq, sel, &nSel, &mean);
      plats[groundInd].quad_status = st;
      if(st)
        info->status |= SSD_STATUS_DEGENERATE_QUAD; /* the reference terminates here */
      else
      {
        plats[groundInd].n_in_quad = nSel;
        plats[groundInd].mean_z = mean;
        if(project_to_binary_image(&f, sel, nSel, img))
          info->status |= SSD_STATUS_BEV_OOB;
        ssd_oracle_close(img, W, H);
        P2 l, r;
        int valid;
        detect_front_edge_closed(img, W, H, &l, &r, &valid);
        P3 *sq = stepQuads[nSteps++];
        if(valid)
        {
          const P2 fl = image_to_world(&f, l), fr = image_to_world(&f, r);
          const LineD frontLine = lined_from_pts(fl, fr);
          const P2 frontLeft = line_intersection_plain(frontLine, lined_from_pts(gq[0], gq[2]));
          const P2 frontRight = line_intersection_plain(frontLine, lined_from_pts(gq[1], gq[3]));
          const P3 a = { frontLeft.x, frontLeft.y, mean }, b = { frontRight.x, frontRight.y, mean }, c = { gq[2].x, gq[2].y, mean },
                   d = { gq[3].x, gq[3].y, mean };
          sq[0] = a;
          sq[1] = b;
          sq[2] = c;
          sq[3] = d;
          if(nSel == 0)
            info->status |= SSD_STATUS_EMPTY_MEAN;
        }
        else
        {
          memset(sq, 0, sizeof(P3) * 4); /* return{} (:546) */
          info->status |= SSD_STATUS_INVALID_FRONT_EDGE;
        }
      }
    }
    for(int i = firstValid; i < nPeaks; i++) /* :445-452, calcStairStep :549-558 */
    {
      if(!plats[i].valid)
        continue;
      uint32_t nSel;
      double mean;
      const int st = points_in_quad_mean(&f, pStart[i], pCount[i], quadW[i], NULL, &nSel, &mean);
      plats[i].quad_status = st;
      if(st)
      {
        info->status |= SSD_STATUS_DEGENERATE_QUAD;
        continue;
      }
      plats[i].n_in_quad = nSel;
      plats[i].mean_z = mean;
      if(nSel == 0)
        info->status |= SSD_STATUS_EMPTY_MEAN;
      P3 *sq = stepQuads[nSteps++];
      for(int c = 0; c < 4; c++)
      {
        sq[c].x = quadW[i][c].x;
        sq[c].y = quadW[i][c].y;
        sq[c].z = mean;
      }
    }
  }
  for(int k = 0; k < nPeaks; k++)
    for(int c = 0; c < 4; c++)
    {
      plats[k].quad_world[c][0] = quadW[k][c].x;
      plats[k].quad_world[c][1] = quadW[k][c].y;
    }

  /* detectStairs result assembly (:370-383) with ToExternalWorld (transformation.cpp:190-194) */
  info->n_steps = nSteps;
  for(int s = 0; s < nSteps; s++)
  {
    for(int c = 0; c < 4; c++)
    {
      g_last_step_quads[s][c][0] = stepQuads[s][c].x;
      g_last_step_quads[s][c][1] = stepQuads[s][c].y;
      g_last_step_quads[s][c][2] = stepQuads[s][c].z;
    }
    for(int c = 0; c < 4; c++)
    {
      const double x = stepQuads[s][c].x, y = stepQuads[s][c].y;
      steps[s].quad[c][0] = (xf->ext_a[0] * x + xf->ext_a[1] * y) + xf->ext_b[0];
      steps[s].quad[c][1] = (xf->ext_a[2] * x + xf->ext_a[3] * y) + xf->ext_b[1];
    }
    steps[s].height = xf->ext_z + stepQuads[s][0].z;
  }
  g_last_n_steps = nSteps;
  if(nSteps == 0)
    info->status |= SSD_STATUS_NO_STEPS;

done:
  free(f.world);
  free(f.pixel);
  free(f.height);
  free(cur);
  free(upper);
  free(store);
  free(sel);
  free(img);
  return rc;
}

/* boost::qvm::inverse of a 3x3 (Transformation_<3>::_aInv = inverse(_a), transformation.cpp:178): adjugate times
 * 1/det, determinant by cofactor expansion along the first row. */
int ssd_oracle_inverse3(const double m[9], double r[9])
{
  const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
  if(det == 0)
    return SSD_E_INVALID_ARG;
  const double f = 1 / det;
  r[0] = f * (m[4] * m[8] - m[5] * m[7]);
  r[1] = f * (m[2] * m[7] - m[1] * m[8]);
  r[2] = f * (m[1] * m[5] - m[2] * m[4]);
  r[3] = f * (m[5] * m[6] - m[3] * m[8]);
  r[4] = f * (m[0] * m[8] - m[2] * m[6]);
  r[5] = f * (m[2] * m[3] - m[0] * m[5]);
  r[6] = f * (m[3] * m[7] - m[4] * m[6]);
  r[7] = f * (m[1] * m[6] - m[0] * m[7]);
  r[8] = f * (m[0] * m[4] - m[1] * m[3]);
  return SSD_OK;
}

/* drawStairStep (pointcloud.cpp:588-597) on the steps of the last ssd_oracle_process of this thread:
 * WorldToCamera = transformInv (transformation.h:66-69: _aInv * (x - _b), QVM mat*vec left to right), narrowing to
 * float and rs2_project_point_to_pixel without distortion (camera.h:80-97). */
int ssd_oracle_last_overlay(const ssd_gpu_transform *xf, const double a_inv[9], const ssd_gpu_intrinsics *intr, ssd_gpu_overlay *out,
                            int cap, int *n)
{
  *n = g_last_n_steps;
  for(int s = 0; s < g_last_n_steps && s < cap; s++)
    for(int c = 0; c < 4; c++)
    {
      const double dx = g_last_step_quads[s][c][0] - xf->b[0], dy = g_last_step_quads[s][c][1] - xf->b[1],
                   dz = g_last_step_quads[s][c][2] - xf->b[2];
      const float X = (float)((a_inv[0] * dx + a_inv[1] * dy) + a_inv[2] * dz);
      const float Y = (float)((a_inv[3] * dx + a_inv[4] * dy) + a_inv[5] * dz);
      const float Z = (float)((a_inv[6] * dx + a_inv[7] * dy) + a_inv[8] * dz);
      const float x = X / Z, y = Y / Z;
      out[s].px[c][0] = x * intr->fx + intr->ppx;
      out[s].px[c][1] = y * intr->fy + intr->ppy;
    }
  return SSD_OK;
}

int ssd_oracle_process_batch(const ssd_gpu_config *cfg, const ssd_gpu_transform *xf, const float *xyz, int n_frames, int *n_steps_out)
{
  const size_t N = (size_t)cfg->width * cfg->height;
  ssd_gpu_frame_info info;
  ssd_gpu_plateau plats[SSD_GPU_MAX_PLATEAUS];
  ssd_gpu_step steps[SSD_GPU_MAX_STEPS];
  for(int i = 0; i < n_frames; i++)
  {
    const int rc = ssd_oracle_process(cfg, xf, xyz + (size_t)i * N * 3, NULL, NULL, 0, &info, plats, steps);
    if(rc)
      return rc;
    if(n_steps_out)
      n_steps_out[i] = info.n_steps;
  }
  return SSD_OK;
}

/* ---------------------------------------------------------------------------------------------
 * Stairs::serialize (stairs.cpp:34-70): fixed, setprecision(3)
 * ------------------------------------------------------------------------------------------- */
int ssd_oracle_serialize(const ssd_gpu_step *steps, int n, char *buf, size_t cap)
{
  size_t len = 0;
#define EMIT(...)                                                                         \
  do                                                                                      \
  {                                                                                       \
    const int w_ = snprintf(buf && len < cap ? buf + len : NULL, buf && len < cap ? cap - len : 0, __VA_ARGS__); \
    len += (size_t)w_;                                                                    \
  } while(0)
  EMIT("[\"stairs\",[\"stairSteps\",%d]", n);
  if(n > 0)
  {
    EMIT(",[");
    for(int i = 0; i < n; i++)
    {
      const ssd_gpu_step *s = &steps[i];
      EMIT("[[\"height\",%.3f],[\"quadrilateral\",[%.3f,%.3f],[%.3f,%.3f],[%.3f,%.3f],[%.3f,%.3f]]]%s", s->height, s->quad[0][0],
           s->quad[0][1], s->quad[1][0], s->quad[1][1], s->quad[2][0], s->quad[2][1], s->quad[3][0], s->quad[3][1], i + 1 < n ? "," : "");
    }
    EMIT("]");
  }
  EMIT("]");
#undef EMIT
  return (int)len;
}
