/*
 * gs_oracle.c -- CPU restatement of the tile-based differentiable 3D-Gaussian-splatting
 * rasterizer that sits behind `diff_gaussian_rasterization` in XuanHuang0/GuassianHand.
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product (guassianhand_b200) never does.
 *
 * PARITY UNPINNED: the reference repo does not contain the rasterizer.  Its arithmetic lives
 * in the un-vendored third-party pip package `diff-gaussian-rasterization==0.0.0`
 * (/root/reference/environment.yml:129; graphdeco-inria/diff-gaussian-rasterization, the
 * original 2-output API identified by the call shape at
 * /root/reference/tgs/models/renderer_one_shot.py:338-346 and the 12-field settings tuple at
 * :281-294).  The reference holds no tests, golden vectors or fixtures for this path
 * (SURVEY.md section 4), so this file restates the *published algorithm* of that package as
 * specified in SURVEY.md Appendix A (A.1-A.8), and is pinned by (i) an independent PyTorch
 * autograd restatement (oracle/torch_ref.py) and (ii) analytic known-answer cases in
 * tests/test_oracle_kat.py.
 *
 * Call sites in the reference that define the inputs this code must accept:
 *   settings tuple     renderer_one_shot.py:281-294, :355-368
 *   forward call       renderer_one_shot.py:338-346 (RGB), :372-379 (mask)
 *   camera matrices    renderer_one_shot.py:61-112 (row-major tensors holding the transpose)
 *
 * Arithmetic contract ("canonical order"): all forward math is IEEE fp32
 * with the fused-multiply-add placement nvcc applies to the natural source order of the
 * published algorithm (verified on nvcc 12.9 SASS):  a*b + c*d + e*f + g  ==
 * ((fma(e,f, fma(a,b, c*d))) + g).   Compile with -ffp-contract=off so gcc adds no fusion of
 * its own; fmaf() is exact-rounded.  The CUDA kernels pin the same order with __fmaf_rn /
 * __fmul_rn / __fadd_rn, so every integer intermediate (radii, tiles, keys, ranges) and every
 * preprocess float is bit-identical.  exp() differs (glibc expf vs libdevice expf, <= 3 ulp),
 * so each pixel carries an `ambig` flag when a threshold decision lies within that band.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GSO_BLOCK 16
#define GSO_EXPORT __attribute__((visibility("default")))

typedef struct {
  int P, H, W, M, D;             /* #gaussians, image, SH coeff count, SH degree */
  float tanfovx, tanfovy, scale_modifier;
  int prefiltered;
  const float *bg;               /* [3] */
  const float *view;             /* [16] flat, element 4*c+r = row r, col c (A.1) */
  const float *proj;             /* [16] */
  const float *campos;           /* [3] */
  const float *means3D;          /* [P,3] */
  const float *opacities;        /* [P] */
  const float *scales;           /* [P,3] or NULL */
  const float *rotations;        /* [P,4] (r,x,y,z) or NULL */
  const float *cov3D_precomp;    /* [P,6] or NULL */
  const float *shs;              /* [P,M,3] or NULL */
  const float *colors_precomp;   /* [P,3] or NULL */
} gso_in;

typedef struct {
  float *depths;        /* [P] */
  int32_t *radii;       /* [P] */
  float *xy;            /* [P,2] */
  float *cov3D;         /* [P,6] */
  float *conic_opacity; /* [P,4] */
  float *rgb;           /* [P,3] */
  uint8_t *clamped;     /* [P,3] */
  uint32_t *tiles_touched; /* [P] */
  uint32_t *offsets;    /* [P] inclusive scan */
  uint64_t R;
  uint64_t *keys_unsorted; /* [R] */
  uint32_t *vals_unsorted; /* [R] */
  uint64_t *keys;       /* [R] sorted */
  uint32_t *point_list; /* [R] sorted */
  uint32_t *ranges;     /* [T,2] */
  float *out_color;     /* [3,H,W] */
  float *final_T;       /* [H*W] */
  uint32_t *n_contrib;  /* [H*W] */
  uint8_t *ambig;       /* [H*W] 1 if an exp-dependent threshold decision is within 4e-6 rel */
  uint64_t n_pairs;     /* sum over pixels of n_contrib */
} gso_fwd;

/* ---- canonical fp32 helpers (A.1: "sums are written left-to-right; default contraction") ---- */
static inline float dot2(float a0, float b0, float a1, float b1) {
  float t = a1 * b1;
  return fmaf(a0, b0, t);
}
static inline float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
  float t = a1 * b1;
  t = fmaf(a0, b0, t);
  return fmaf(a2, b2, t);
}
static inline float dot3a(float a0, float b0, float a1, float b1, float a2, float b2, float c) {
  return dot3(a0, b0, a1, b1, a2, b2) + c;
}
static inline float fminf_(float a, float b) { return a < b ? a : b; }
static inline float fmaxf_(float a, float b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

/* A.2 step 3: Sigma = R S S^T R^T, A = R diag(s) */
static void cov3d_from_scale_rot(const float *scale, float mod, const float *q, float *cov6) {
  float s0 = mod * scale[0], s1 = mod * scale[1], s2 = mod * scale[2];
  float r = q[0], x = q[1], y = q[2], z = q[3];
  float R[3][3];
  R[0][0] = fmaf(-2.f, fmaf(y, y, z * z), 1.f);
  R[0][1] = 2.f * fmaf(x, y, -(r * z));
  R[0][2] = 2.f * fmaf(x, z, r * y);
  R[1][0] = 2.f * fmaf(x, y, r * z);
  R[1][1] = fmaf(-2.f, fmaf(x, x, z * z), 1.f);
  R[1][2] = 2.f * fmaf(y, z, -(r * x));
  R[2][0] = 2.f * fmaf(x, z, -(r * y));
  R[2][1] = 2.f * fmaf(y, z, r * x);
  R[2][2] = fmaf(-2.f, fmaf(x, x, y * y), 1.f);
  float A[3][3];
  for (int i = 0; i < 3; i++) {
    A[i][0] = R[i][0] * s0;
    A[i][1] = R[i][1] * s1;
    A[i][2] = R[i][2] * s2;
  }
  cov6[0] = dot3(A[0][0], A[0][0], A[0][1], A[0][1], A[0][2], A[0][2]);
  cov6[1] = dot3(A[0][0], A[1][0], A[0][1], A[1][1], A[0][2], A[1][2]);
  cov6[2] = dot3(A[0][0], A[2][0], A[0][1], A[2][1], A[0][2], A[2][2]);
  cov6[3] = dot3(A[1][0], A[1][0], A[1][1], A[1][1], A[1][2], A[1][2]);
  cov6[4] = dot3(A[1][0], A[2][0], A[1][1], A[2][1], A[1][2], A[2][2]);
  cov6[5] = dot3(A[2][0], A[2][0], A[2][1], A[2][1], A[2][2], A[2][2]);
}

typedef struct {
  float T[2][3];   /* J*W rows 0,1 */
  float tx, ty, tz;/* clamped view-space point */
  float gx, gy;    /* clamp gradient gates */
  float fx, fy;
} cov2d_ctx;

/* A.2 step 4 */
static void cov2d(const float *view, float tanfovx, float tanfovy, int W, int H, float pvx, float pvy,
                  float pvz, const float *c3, float *a, float *b, float *c, cov2d_ctx *ctx) {
  float fx = (float)W / (2.0f * tanfovx);
  float fy = (float)H / (2.0f * tanfovy);
  float limx = 1.3f * tanfovx, limy = 1.3f * tanfovy;
  float txtz = pvx / pvz, tytz = pvy / pvz;
  float tx = fminf_(limx, fmaxf_(-limx, txtz)) * pvz;
  float ty = fminf_(limy, fmaxf_(-limy, tytz)) * pvz;
  float tz = pvz;
  float J00 = fx / tz;
  float J02 = -(fx * tx) / (tz * tz);
  float J11 = fy / tz;
  float J12 = -(fy * ty) / (tz * tz);
  float T0[3], T1[3];
  for (int cc = 0; cc < 3; cc++) {
    float W0 = view[4 * cc + 0], W1 = view[4 * cc + 1], W2 = view[4 * cc + 2];
    T0[cc] = dot2(J00, W0, J02, W2);
    T1[cc] = dot2(J11, W1, J12, W2);
  }
  /* u_r = Sigma * T_r */
  float S[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
  float u0[3], u1[3];
  for (int i = 0; i < 3; i++) {
    u0[i] = dot3(S[i][0], T0[0], S[i][1], T0[1], S[i][2], T0[2]);
    u1[i] = dot3(S[i][0], T1[0], S[i][1], T1[1], S[i][2], T1[2]);
  }
  *a = dot3(T0[0], u0[0], T0[1], u0[1], T0[2], u0[2]) + 0.3f;
  *b = dot3(T0[0], u1[0], T0[1], u1[1], T0[2], u1[2]);
  *c = dot3(T1[0], u1[0], T1[1], u1[1], T1[2], u1[2]) + 0.3f;
  if (ctx) {
    for (int i = 0; i < 3; i++) { ctx->T[0][i] = T0[i]; ctx->T[1][i] = T1[i]; }
    ctx->tx = tx; ctx->ty = ty; ctx->tz = tz;
    ctx->gx = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    ctx->gy = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    ctx->fx = fx; ctx->fy = fy;
  }
}

/* A.3 */
static void sh_to_rgb(int D, int M, const float *sh /*[M,3]*/, const float *p, const float *campos,
                      float *rgb, uint8_t *clamped) {
  (void)M;
  float dx = p[0] - campos[0], dy = p[1] - campos[1], dz = p[2] - campos[2];
  float len = sqrtf(dot3(dx, dx, dy, dy, dz, dz));
  float x = dx / len, y = dy / len, z = dz / len;
  for (int ch = 0; ch < 3; ch++) {
    float r = SH_C0 * sh[0 * 3 + ch];
    if (D > 0) {
      r = fmaf(-(SH_C1 * y), sh[1 * 3 + ch], r);
      r = fmaf(SH_C1 * z, sh[2 * 3 + ch], r);
      r = fmaf(-(SH_C1 * x), sh[3 * 3 + ch], r);
      if (D > 1) {
        float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        r = fmaf(SH_C2[0] * xy, sh[4 * 3 + ch], r);
        r = fmaf(SH_C2[1] * yz, sh[5 * 3 + ch], r);
        r = fmaf(SH_C2[2] * (fmaf(2.0f, zz, -xx) - yy), sh[6 * 3 + ch], r);
        r = fmaf(SH_C2[3] * xz, sh[7 * 3 + ch], r);
        r = fmaf(SH_C2[4] * (xx - yy), sh[8 * 3 + ch], r);
        if (D > 2) {
          r = fmaf(SH_C3[0] * y * (fmaf(3.0f, xx, -yy)), sh[9 * 3 + ch], r);
          r = fmaf(SH_C3[1] * xy * z, sh[10 * 3 + ch], r);
          r = fmaf(SH_C3[2] * y * (fmaf(4.0f, zz, -xx) - yy), sh[11 * 3 + ch], r);
          r = fmaf(SH_C3[3] * z * (fmaf(-3.0f, yy, fmaf(2.0f, zz, -(3.0f * xx)))), sh[12 * 3 + ch], r);
          r = fmaf(SH_C3[4] * x * (fmaf(4.0f, zz, -xx) - yy), sh[13 * 3 + ch], r);
          r = fmaf(SH_C3[5] * z * (xx - yy), sh[14 * 3 + ch], r);
          r = fmaf(SH_C3[6] * x * (fmaf(-3.0f, yy, xx)), sh[15 * 3 + ch], r);
        }
      }
    }
    r += 0.5f;
    clamped[ch] = (r < 0.f);
    rgb[ch] = fmaxf_(r, 0.f);
  }
}

static void get_rect(float px, float py, int radius, int gx, int gy, int *minx, int *miny, int *maxx,
                     int *maxy) {
  float r = (float)radius;
  *minx = imin(gx, imax(0, (int)((px - r) / 16.0f)));
  *miny = imin(gy, imax(0, (int)((py - r) / 16.0f)));
  *maxx = imin(gx, imax(0, (int)((((px + r) + 16.0f) - 1.0f) / 16.0f)));
  *maxy = imin(gy, imax(0, (int)((((py + r) + 16.0f) - 1.0f) / 16.0f)));
}

/* stable LSD radix sort of (key,val) pairs on the low `nbits` bits (A.4) */
static void radix_sort_pairs(uint64_t *keys, uint32_t *vals, uint64_t *keys_tmp, uint32_t *vals_tmp,
                             uint64_t n, int nbits) {
  uint64_t *kin = keys, *kout = keys_tmp;
  uint32_t *vin = vals, *vout = vals_tmp;
  int passes = (nbits + 15) / 16;
  size_t *hist = (size_t *)malloc(sizeof(size_t) * 65536);
  for (int p = 0; p < passes; p++) {
    int shift = 16 * p;
    memset(hist, 0, sizeof(size_t) * 65536);
    for (uint64_t i = 0; i < n; i++) hist[(kin[i] >> shift) & 0xFFFF]++;
    size_t sum = 0;
    for (int d = 0; d < 65536; d++) { size_t c = hist[d]; hist[d] = sum; sum += c; }
    for (uint64_t i = 0; i < n; i++) {
      size_t pos = hist[(kin[i] >> shift) & 0xFFFF]++;
      kout[pos] = kin[i];
      vout[pos] = vin[i];
    }
    uint64_t *tk = kin; kin = kout; kout = tk;
    uint32_t *tv = vin; vin = vout; vout = tv;
  }
  if (kin != keys) {
    memcpy(keys, kin, n * sizeof(uint64_t));
    memcpy(vals, vin, n * sizeof(uint32_t));
  }
  free(hist);
}

static int higher_msb(uint32_t n) { /* A.4: bisection starting at 16 */
  uint32_t msb = 16, step = 16;
  while (step > 1) {
    step /= 2;
    if (n >> msb) msb += step; else msb -= step;
  }
  if (n >> msb) msb++;
  return (int)msb;
}

GSO_EXPORT void gso_free(gso_fwd *f) {
  if (!f) return;
  free(f->depths); free(f->radii); free(f->xy); free(f->cov3D); free(f->conic_opacity); free(f->rgb);
  free(f->clamped); free(f->tiles_touched); free(f->offsets); free(f->keys_unsorted);
  free(f->vals_unsorted); free(f->keys); free(f->point_list); free(f->ranges); free(f->out_color);
  free(f->final_T); free(f->n_contrib); free(f->ambig);
  free(f);
}

GSO_EXPORT int gso_higher_msb(uint32_t n) { return higher_msb(n); }

/* A.2: per-Gaussian preprocess.  Returns 1 if visible. */
static int preprocess_one(const gso_in *in, int i, gso_fwd *f, int gx, int gy) {
  const float *p = in->means3D + 3 * i;
  const float *V = in->view, *PV = in->proj;
  f->radii[i] = 0;
  f->tiles_touched[i] = 0;
  float pvx = dot3a(V[0], p[0], V[4], p[1], V[8], p[2], V[12]);
  float pvy = dot3a(V[1], p[0], V[5], p[1], V[9], p[2], V[13]);
  float pvz = dot3a(V[2], p[0], V[6], p[1], V[10], p[2], V[14]);
  if (pvz <= 0.2f) return 0;
  float phx = dot3a(PV[0], p[0], PV[4], p[1], PV[8], p[2], PV[12]);
  float phy = dot3a(PV[1], p[0], PV[5], p[1], PV[9], p[2], PV[13]);
  float phw = dot3a(PV[3], p[0], PV[7], p[1], PV[11], p[2], PV[15]);
  float p_w = 1.0f / (phw + 0.0000001f);
  float ppx = phx * p_w, ppy = phy * p_w;
  float *c3 = f->cov3D + 6 * i;
  if (in->cov3D_precomp) memcpy(c3, in->cov3D_precomp + 6 * i, 6 * sizeof(float));
  else cov3d_from_scale_rot(in->scales + 3 * i, in->scale_modifier, in->rotations + 4 * i, c3);
  float a, b, c;
  cov2d(V, in->tanfovx, in->tanfovy, in->W, in->H, pvx, pvy, pvz, c3, &a, &b, &c, NULL);
  float det = fmaf(a, c, -(b * b));
  if (det == 0.0f) return 0;
  float det_inv = 1.f / det;
  float conx = c * det_inv, cony = -b * det_inv, conz = a * det_inv;
  float mid = 0.5f * (a + c);
  float sq = sqrtf(fmaxf_(0.1f, fmaf(mid, mid, -det)));
  float l1 = mid + sq, l2 = mid - sq;
  int radius = (int)ceilf(3.f * sqrtf(fmaxf_(l1, l2)));
  float pix_x = fmaf(ppx + 1.0f, (float)in->W, -1.0f) * 0.5f;
  float pix_y = fmaf(ppy + 1.0f, (float)in->H, -1.0f) * 0.5f;
  int minx, miny, maxx, maxy;
  get_rect(pix_x, pix_y, radius, gx, gy, &minx, &miny, &maxx, &maxy);
  if ((maxx - minx) * (maxy - miny) == 0) return 0;
  if (in->colors_precomp) {
    memcpy(f->rgb + 3 * i, in->colors_precomp + 3 * i, 3 * sizeof(float));
  } else {
    sh_to_rgb(in->D, in->M, in->shs + (size_t)i * in->M * 3, p, in->campos, f->rgb + 3 * i,
              f->clamped + 3 * i);
  }
  f->depths[i] = pvz;
  f->radii[i] = radius;
  f->xy[2 * i] = pix_x;
  f->xy[2 * i + 1] = pix_y;
  f->conic_opacity[4 * i + 0] = conx;
  f->conic_opacity[4 * i + 1] = cony;
  f->conic_opacity[4 * i + 2] = conz;
  f->conic_opacity[4 * i + 3] = in->opacities[i];
  f->tiles_touched[i] = (uint32_t)((maxx - minx) * (maxy - miny));
  return 1;
}

/* A.5: one tile */
static void render_tile(const gso_in *in, const gso_fwd *f, int tx, int ty, int gx, uint64_t *pairs) {
  int W = in->W, H = in->H;
  uint32_t start = f->ranges[2 * (ty * gx + tx)], end = f->ranges[2 * (ty * gx + tx) + 1];
  const float amin = 1.0f / 255.0f;
  for (int ly = 0; ly < GSO_BLOCK; ly++) {
    int py = ty * GSO_BLOCK + ly;
    if (py >= H) break;
    for (int lx = 0; lx < GSO_BLOCK; lx++) {
      int px = tx * GSO_BLOCK + lx;
      if (px >= W) break;
      float pxf = (float)px, pyf = (float)py;
      float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
      uint32_t contributor = 0, last = 0;
      uint8_t amb = 0;
      for (uint32_t j = start; j < end; j++) {
        contributor++;
        uint32_t id = f->point_list[j];
        float dx = f->xy[2 * id] - pxf, dy = f->xy[2 * id + 1] - pyf;
        const float *co = f->conic_opacity + 4 * id;
        float q = fmaf(co[0] * dx, dx, (co[2] * dy) * dy);
        float power = fmaf(-0.5f, q, -((co[1] * dx) * dy));
        if (power > 0.0f) continue;
        float oa = co[3] * expf(power);
        float alpha = fminf_(0.99f, oa);
        if (fabsf(oa - amin) <= 4e-6f * amin) amb = 1;
        if (alpha < amin) continue;
        float test_T = T * (1.f - alpha);
        if (fabsf(test_T - 0.0001f) <= 2e-5f * 0.0001f) amb = 1;
        if (test_T < 0.0001f) break;
        const float *col = f->rgb + 3 * id;
        C0 = fmaf(col[0] * alpha, T, C0);
        C1 = fmaf(col[1] * alpha, T, C1);
        C2 = fmaf(col[2] * alpha, T, C2);
        T = test_T;
        last = contributor;
      }
      size_t pix = (size_t)py * W + px;
      f->final_T[pix] = T;
      f->n_contrib[pix] = last;
      f->ambig[pix] = amb;
      f->out_color[0 * (size_t)H * W + pix] = fmaf(T, in->bg[0], C0);
      f->out_color[1 * (size_t)H * W + pix] = fmaf(T, in->bg[1], C1);
      f->out_color[2 * (size_t)H * W + pix] = fmaf(T, in->bg[2], C2);
      *pairs += last;
    }
  }
}

GSO_EXPORT gso_fwd *gso_forward(const gso_in *in) {
  int P = in->P, H = in->H, W = in->W;
  int gx = (W + GSO_BLOCK - 1) / GSO_BLOCK, gy = (H + GSO_BLOCK - 1) / GSO_BLOCK;
  int Tn = gx * gy;
  size_t N = (size_t)H * W;
  gso_fwd *f = (gso_fwd *)calloc(1, sizeof(gso_fwd));
  size_t Pa = P > 0 ? (size_t)P : 1;
  f->depths = (float *)calloc(Pa, sizeof(float));
  f->radii = (int32_t *)calloc(Pa, sizeof(int32_t));
  f->xy = (float *)calloc(Pa * 2, sizeof(float));
  f->cov3D = (float *)calloc(Pa * 6, sizeof(float));
  f->conic_opacity = (float *)calloc(Pa * 4, sizeof(float));
  f->rgb = (float *)calloc(Pa * 3, sizeof(float));
  f->clamped = (uint8_t *)calloc(Pa * 3, 1);
  f->tiles_touched = (uint32_t *)calloc(Pa, sizeof(uint32_t));
  f->offsets = (uint32_t *)calloc(Pa, sizeof(uint32_t));
  f->ranges = (uint32_t *)calloc((size_t)(Tn > 0 ? Tn : 1) * 2, sizeof(uint32_t));
  f->out_color = (float *)calloc((N ? N : 1) * 3, sizeof(float));
  f->final_T = (float *)calloc(N ? N : 1, sizeof(float));
  f->n_contrib = (uint32_t *)calloc(N ? N : 1, sizeof(uint32_t));
  f->ambig = (uint8_t *)calloc(N ? N : 1, 1);

#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) preprocess_one(in, i, f, gx, gy);

  /* A.4 */
  uint64_t sum = 0;
  for (int i = 0; i < P; i++) { sum += f->tiles_touched[i]; f->offsets[i] = (uint32_t)sum; }
  uint64_t R = sum;
  f->R = R;
  size_t Ra = R ? R : 1;
  f->keys_unsorted = (uint64_t *)malloc(Ra * sizeof(uint64_t));
  f->vals_unsorted = (uint32_t *)malloc(Ra * sizeof(uint32_t));
  f->keys = (uint64_t *)malloc(Ra * sizeof(uint64_t));
  f->point_list = (uint32_t *)malloc(Ra * sizeof(uint32_t));
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    if (f->radii[i] > 0) {
      uint32_t off = (i == 0) ? 0 : f->offsets[i - 1];
      int minx, miny, maxx, maxy;
      get_rect(f->xy[2 * i], f->xy[2 * i + 1], f->radii[i], gx, gy, &minx, &miny, &maxx, &maxy);
      uint32_t dbits;
      memcpy(&dbits, &f->depths[i], 4);
      for (int y = miny; y < maxy; y++)
        for (int x = minx; x < maxx; x++) {
          uint64_t key = (uint64_t)(uint32_t)(y * gx + x);
          key <<= 32;
          key |= dbits;
          f->keys_unsorted[off] = key;
          f->vals_unsorted[off] = (uint32_t)i;
          off++;
        }
    }
  }
  memcpy(f->keys, f->keys_unsorted, R * sizeof(uint64_t));
  memcpy(f->point_list, f->vals_unsorted, R * sizeof(uint32_t));
  {
    uint64_t *kt = (uint64_t *)malloc(Ra * sizeof(uint64_t));
    uint32_t *vt = (uint32_t *)malloc(Ra * sizeof(uint32_t));
    int bit = higher_msb((uint32_t)Tn);
    radix_sort_pairs(f->keys, f->point_list, kt, vt, R, 32 + bit);
    free(kt); free(vt);
  }
  for (uint64_t idx = 0; idx < R; idx++) {
    uint32_t cur = (uint32_t)(f->keys[idx] >> 32);
    if (idx == 0) f->ranges[2 * cur] = 0;
    else {
      uint32_t prev = (uint32_t)(f->keys[idx - 1] >> 32);
      if (cur != prev) { f->ranges[2 * prev + 1] = (uint32_t)idx; f->ranges[2 * cur] = (uint32_t)idx; }
    }
    if (idx == R - 1) f->ranges[2 * cur + 1] = (uint32_t)R;
  }

  uint64_t pairs = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : pairs)
  for (int t = 0; t < Tn; t++) {
    uint64_t pp = 0;
    render_tile(in, f, t % gx, t / gx, gx, &pp);
    pairs += pp;
  }
  f->n_pairs = pairs;
  return f;
}

/* ------------------------------------------------------------------ backward */

static inline void atomic_addd(double *p, double v) {
#pragma omp atomic
  *p += v;
}

/* A.6: one tile; accumulators are double so the sum order does not matter at fp32 level */
static void render_tile_bwd(const gso_in *in, const gso_fwd *f, const float *dL_dout, int tx, int ty, int gx,
                            double *dmean2D /*[P,2]*/, double *dconic /*[P,3]*/, double *dopac /*[P]*/,
                            double *dcolor /*[P,3]*/) {
  int W = in->W, H = in->H;
  size_t N = (size_t)H * W;
  uint32_t start = f->ranges[2 * (ty * gx + tx)], end = f->ranges[2 * (ty * gx + tx) + 1];
  const float amin = 1.0f / 255.0f;
  const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
  for (int ly = 0; ly < GSO_BLOCK; ly++) {
    int py = ty * GSO_BLOCK + ly;
    if (py >= H) break;
    for (int lx = 0; lx < GSO_BLOCK; lx++) {
      int px = tx * GSO_BLOCK + lx;
      if (px >= W) break;
      size_t pix = (size_t)py * W + px;
      float pxf = (float)px, pyf = (float)py;
      float T_final = f->final_T[pix];
      float T = T_final;
      uint32_t last = f->n_contrib[pix];
      float dLp[3] = {dL_dout[pix], dL_dout[N + pix], dL_dout[2 * N + pix]};
      float accum[3] = {0, 0, 0}, last_c[3] = {0, 0, 0};
      float last_alpha = 0.f;
      float bgdot = 0.f;
      for (int ch = 0; ch < 3; ch++) bgdot += in->bg[ch] * dLp[ch];
      for (uint32_t k = 0; k < last && k < end - start; k++) {
        uint32_t j = start + (last - 1 - k);
        uint32_t id = f->point_list[j];
        float dx = f->xy[2 * id] - pxf, dy = f->xy[2 * id + 1] - pyf;
        const float *co = f->conic_opacity + 4 * id;
        float q = fmaf(co[0] * dx, dx, (co[2] * dy) * dy);
        float power = fmaf(-0.5f, q, -((co[1] * dx) * dy));
        if (power > 0.0f) continue;
        float G = expf(power);
        float alpha = fminf_(0.99f, co[3] * G);
        if (alpha < amin) continue;
        T = T / (1.f - alpha);
        float dchannel_dcolor = alpha * T;
        float dL_dalpha = 0.f;
        const float *col = f->rgb + 3 * id;
        for (int ch = 0; ch < 3; ch++) {
          accum[ch] = last_alpha * last_c[ch] + (1.f - last_alpha) * accum[ch];
          last_c[ch] = col[ch];
          dL_dalpha += (col[ch] - accum[ch]) * dLp[ch];
          atomic_addd(&dcolor[3 * id + ch], (double)(dchannel_dcolor * dLp[ch]));
        }
        dL_dalpha *= T;
        last_alpha = alpha;
        dL_dalpha += (-T_final / (1.f - alpha)) * bgdot;
        float dL_dG = co[3] * dL_dalpha;
        float gdx = G * dx, gdy = G * dy;
        float dG_ddelx = -gdx * co[0] - gdy * co[1];
        float dG_ddely = -gdy * co[2] - gdx * co[1];
        atomic_addd(&dmean2D[2 * id + 0], (double)(dL_dG * dG_ddelx * ddelx_dx));
        atomic_addd(&dmean2D[2 * id + 1], (double)(dL_dG * dG_ddely * ddely_dy));
        atomic_addd(&dconic[3 * id + 0], (double)(-0.5f * gdx * dx * dL_dG));
        atomic_addd(&dconic[3 * id + 1], (double)(-0.5f * gdx * dy * dL_dG));
        atomic_addd(&dconic[3 * id + 2], (double)(-0.5f * gdy * dy * dL_dG));
        atomic_addd(&dopac[id], (double)(G * dL_dalpha));
      }
    }
  }
}

/* outputs are caller-allocated and fully overwritten */
GSO_EXPORT int gso_backward(const gso_in *in, const gso_fwd *f, const float *dL_dout,
                            float *dL_dmeans3D /*[P,3]*/, float *dL_dmeans2D /*[P,3]*/,
                            float *dL_dcolors /*[P,3]*/, float *dL_dconic /*[P,4] (xx,xy,-,yy)*/,
                            float *dL_dopacity /*[P]*/, float *dL_dcov3D /*[P,6]*/,
                            float *dL_dsh /*[P,M,3] or NULL*/, float *dL_dscales /*[P,3]*/,
                            float *dL_drots /*[P,4]*/) {
  int P = in->P, H = in->H, W = in->W, M = in->M;
  int gx = (W + GSO_BLOCK - 1) / GSO_BLOCK, gy = (H + GSO_BLOCK - 1) / GSO_BLOCK;
  int Tn = gx * gy;
  size_t Pa = P > 0 ? (size_t)P : 1;
  double *dmean2D = (double *)calloc(Pa * 2, sizeof(double));
  double *dconic = (double *)calloc(Pa * 3, sizeof(double));
  double *dopac = (double *)calloc(Pa, sizeof(double));
  double *dcolor = (double *)calloc(Pa * 3, sizeof(double));
#pragma omp parallel for schedule(dynamic, 1)
  for (int t = 0; t < Tn; t++)
    render_tile_bwd(in, f, dL_dout, t % gx, t / gx, gx, dmean2D, dconic, dopac, dcolor);

  memset(dL_dmeans3D, 0, Pa * 3 * sizeof(float));
  memset(dL_dmeans2D, 0, Pa * 3 * sizeof(float));
  memset(dL_dcolors, 0, Pa * 3 * sizeof(float));
  memset(dL_dconic, 0, Pa * 4 * sizeof(float));
  memset(dL_dopacity, 0, Pa * sizeof(float));
  memset(dL_dcov3D, 0, Pa * 6 * sizeof(float));
  if (dL_dsh && M > 0) memset(dL_dsh, 0, Pa * (size_t)M * 3 * sizeof(float));
  memset(dL_dscales, 0, Pa * 3 * sizeof(float));
  memset(dL_drots, 0, Pa * 4 * sizeof(float));

  const float *V = in->view, *PV = in->proj;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    dL_dmeans2D[3 * i + 0] = (float)dmean2D[2 * i + 0];
    dL_dmeans2D[3 * i + 1] = (float)dmean2D[2 * i + 1];
    dL_dconic[4 * i + 0] = (float)dconic[3 * i + 0];
    dL_dconic[4 * i + 1] = (float)dconic[3 * i + 1];
    dL_dconic[4 * i + 3] = (float)dconic[3 * i + 2];
    dL_dopacity[i] = (float)dopac[i];
    for (int ch = 0; ch < 3; ch++) dL_dcolors[3 * i + ch] = (float)dcolor[3 * i + ch];
    if (!(f->radii[i] > 0)) continue;
    const float *p = in->means3D + 3 * i;
    float gA = dL_dconic[4 * i + 0], gB = dL_dconic[4 * i + 1], gC = dL_dconic[4 * i + 3];
    /* ---- A.7 cov2D backward ---- */
    float pvx = dot3a(V[0], p[0], V[4], p[1], V[8], p[2], V[12]);
    float pvy = dot3a(V[1], p[0], V[5], p[1], V[9], p[2], V[13]);
    float pvz = dot3a(V[2], p[0], V[6], p[1], V[10], p[2], V[14]);
    const float *c3 = f->cov3D + 6 * i;
    float a, b, c;
    cov2d_ctx cx;
    cov2d(V, in->tanfovx, in->tanfovy, W, H, pvx, pvy, pvz, c3, &a, &b, &c, &cx);
    float denom = fmaf(a, c, -(b * b));   /* same rounding as the forward's det */
    float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    float dmx = 0.f, dmy = 0.f, dmz = 0.f;
    float dc3[6] = {0, 0, 0, 0, 0, 0};
    if (denom2inv != 0) {
      float dL_da = denom2inv * (-c * c * gA + 2 * b * c * gB + (denom - a * c) * gC);
      float dL_dc = denom2inv * (-a * a * gC + 2 * a * b * gB + (denom - a * c) * gA);
      float dL_db = denom2inv * 2 * (b * c * gA - (denom + 2 * b * b) * gB + a * b * gC);
      const float *T0 = cx.T[0], *T1 = cx.T[1];
      dc3[0] = T0[0] * T0[0] * dL_da + T0[0] * T1[0] * dL_db + T1[0] * T1[0] * dL_dc;
      dc3[3] = T0[1] * T0[1] * dL_da + T0[1] * T1[1] * dL_db + T1[1] * T1[1] * dL_dc;
      dc3[5] = T0[2] * T0[2] * dL_da + T0[2] * T1[2] * dL_db + T1[2] * T1[2] * dL_dc;
      dc3[1] = 2 * T0[0] * T0[1] * dL_da + (T0[0] * T1[1] + T0[1] * T1[0]) * dL_db + 2 * T1[0] * T1[1] * dL_dc;
      dc3[2] = 2 * T0[0] * T0[2] * dL_da + (T0[0] * T1[2] + T0[2] * T1[0]) * dL_db + 2 * T1[0] * T1[2] * dL_dc;
      dc3[4] = 2 * T0[2] * T0[1] * dL_da + (T0[1] * T1[2] + T0[2] * T1[1]) * dL_db + 2 * T1[1] * T1[2] * dL_dc;
      float S[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
      float u0[3], u1[3];
      for (int r = 0; r < 3; r++) {
        u0[r] = S[r][0] * T0[0] + S[r][1] * T0[1] + S[r][2] * T0[2];
        u1[r] = S[r][0] * T1[0] + S[r][1] * T1[1] + S[r][2] * T1[2];
      }
      float dT0[3], dT1[3];
      for (int r = 0; r < 3; r++) {
        dT0[r] = 2 * u0[r] * dL_da + u1[r] * dL_db;
        dT1[r] = 2 * u1[r] * dL_dc + u0[r] * dL_db;
      }
      float dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
      for (int r = 0; r < 3; r++) {
        float W0 = V[4 * r + 0], W1 = V[4 * r + 1], W2 = V[4 * r + 2];
        dJ00 += W0 * dT0[r];
        dJ02 += W2 * dT0[r];
        dJ11 += W1 * dT1[r];
        dJ12 += W2 * dT1[r];
      }
      float tz = 1.f / cx.tz, tz2 = tz * tz, tz3 = tz2 * tz;
      float dtx = cx.gx * -cx.fx * tz2 * dJ02;
      float dty = cx.gy * -cx.fy * tz2 * dJ12;
      float dtz = -cx.fx * tz2 * dJ00 - cx.fy * tz2 * dJ11 + (2 * cx.fx * cx.tx) * tz3 * dJ02 +
                  (2 * cx.fy * cx.ty) * tz3 * dJ12;
      dmx = V[0] * dtx + V[1] * dty + V[2] * dtz;
      dmy = V[4] * dtx + V[5] * dty + V[6] * dtz;
      dmz = V[8] * dtx + V[9] * dty + V[10] * dtz;
    }
    for (int k = 0; k < 6; k++) dL_dcov3D[6 * i + k] = dc3[k];
    /* ---- A.8 projection backward ---- */
    float phw = dot3a(PV[3], p[0], PV[7], p[1], PV[11], p[2], PV[15]);
    float m_w = 1.0f / (phw + 0.0000001f);
    float mul1 = (PV[0] * p[0] + PV[4] * p[1] + PV[8] * p[2] + PV[12]) * m_w * m_w;
    float mul2 = (PV[1] * p[0] + PV[5] * p[1] + PV[9] * p[2] + PV[13]) * m_w * m_w;
    float g2x = dL_dmeans2D[3 * i + 0], g2y = dL_dmeans2D[3 * i + 1];
    dmx += (PV[0] * m_w - PV[3] * mul1) * g2x + (PV[1] * m_w - PV[3] * mul2) * g2y;
    dmy += (PV[4] * m_w - PV[7] * mul1) * g2x + (PV[5] * m_w - PV[7] * mul2) * g2y;
    dmz += (PV[8] * m_w - PV[11] * mul1) * g2x + (PV[9] * m_w - PV[11] * mul2) * g2y;
    /* ---- A.3 SH backward ---- */
    if (in->shs && !in->colors_precomp && dL_dsh) {
      int D = in->D;
      const float *sh = in->shs + (size_t)i * M * 3;
      float *dsh = dL_dsh + (size_t)i * M * 3;
      float ox = p[0] - in->campos[0], oy = p[1] - in->campos[1], oz = p[2] - in->campos[2];
      float len = sqrtf(ox * ox + oy * oy + oz * oz);
      float x = ox / len, y = oy / len, z = oz / len;
      float dRGB[3];
      for (int ch = 0; ch < 3; ch++) dRGB[ch] = f->clamped[3 * i + ch] ? 0.f : dL_dcolors[3 * i + ch];
      float dRdx[3] = {0, 0, 0}, dRdy[3] = {0, 0, 0}, dRdz[3] = {0, 0, 0};
      for (int ch = 0; ch < 3; ch++) {
        float g = dRGB[ch];
#define SH(k) sh[(k) * 3 + ch]
#define DSH(k) dsh[(k) * 3 + ch]
        DSH(0) = SH_C0 * g;
        if (D > 0) {
          DSH(1) = -SH_C1 * y * g;
          DSH(2) = SH_C1 * z * g;
          DSH(3) = -SH_C1 * x * g;
          dRdx[ch] = -SH_C1 * SH(3);
          dRdy[ch] = -SH_C1 * SH(1);
          dRdz[ch] = SH_C1 * SH(2);
          if (D > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            DSH(4) = SH_C2[0] * xy * g;
            DSH(5) = SH_C2[1] * yz * g;
            DSH(6) = SH_C2[2] * (2.f * zz - xx - yy) * g;
            DSH(7) = SH_C2[3] * xz * g;
            DSH(8) = SH_C2[4] * (xx - yy) * g;
            dRdx[ch] += SH_C2[0] * y * SH(4) + SH_C2[2] * 2.f * -x * SH(6) + SH_C2[3] * z * SH(7) +
                        SH_C2[4] * 2.f * x * SH(8);
            dRdy[ch] += SH_C2[0] * x * SH(4) + SH_C2[1] * z * SH(5) + SH_C2[2] * 2.f * -y * SH(6) +
                        SH_C2[4] * 2.f * -y * SH(8);
            dRdz[ch] += SH_C2[1] * y * SH(5) + SH_C2[2] * 2.f * 2.f * z * SH(6) + SH_C2[3] * x * SH(7);
            if (D > 2) {
              DSH(9) = SH_C3[0] * y * (3.f * xx - yy) * g;
              DSH(10) = SH_C3[1] * xy * z * g;
              DSH(11) = SH_C3[2] * y * (4.f * zz - xx - yy) * g;
              DSH(12) = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy) * g;
              DSH(13) = SH_C3[4] * x * (4.f * zz - xx - yy) * g;
              DSH(14) = SH_C3[5] * z * (xx - yy) * g;
              DSH(15) = SH_C3[6] * x * (xx - 3.f * yy) * g;
              dRdx[ch] += SH_C3[0] * SH(9) * 3.f * 2.f * xy + SH_C3[1] * SH(10) * yz +
                          SH_C3[2] * SH(11) * -2.f * xy + SH_C3[3] * SH(12) * -3.f * 2.f * xz +
                          SH_C3[4] * SH(13) * (-3.f * xx + 4.f * zz - yy) + SH_C3[5] * SH(14) * 2.f * xz +
                          SH_C3[6] * SH(15) * 3.f * (xx - yy);
              dRdy[ch] += SH_C3[0] * SH(9) * 3.f * (xx - yy) + SH_C3[1] * SH(10) * xz +
                          SH_C3[2] * SH(11) * (-3.f * yy + 4.f * zz - xx) + SH_C3[3] * SH(12) * -3.f * 2.f * yz +
                          SH_C3[4] * SH(13) * -2.f * xy + SH_C3[5] * SH(14) * -2.f * yz +
                          SH_C3[6] * SH(15) * -3.f * 2.f * xy;
              dRdz[ch] += SH_C3[1] * SH(10) * xy + SH_C3[2] * SH(11) * 4.f * 2.f * yz +
                          SH_C3[3] * SH(12) * 3.f * (2.f * zz - xx - yy) + SH_C3[4] * SH(13) * 4.f * 2.f * xz +
                          SH_C3[5] * SH(14) * (xx - yy);
            }
          }
        }
#undef SH
#undef DSH
      }
      float ddx = dRdx[0] * dRGB[0] + dRdx[1] * dRGB[1] + dRdx[2] * dRGB[2];
      float ddy = dRdy[0] * dRGB[0] + dRdy[1] * dRGB[1] + dRdy[2] * dRGB[2];
      float ddz = dRdz[0] * dRGB[0] + dRdz[1] * dRGB[1] + dRdz[2] * dRGB[2];
      float sum2 = ox * ox + oy * oy + oz * oz;
      float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
      dmx += ((sum2 - ox * ox) * ddx - oy * ox * ddy - oz * ox * ddz) * invsum32;
      dmy += (-ox * oy * ddx + (sum2 - oy * oy) * ddy - oz * oy * ddz) * invsum32;
      dmz += (-ox * oz * ddx - oy * oz * ddy + (sum2 - oz * oz) * ddz) * invsum32;
    }
    dL_dmeans3D[3 * i + 0] = dmx;
    dL_dmeans3D[3 * i + 1] = dmy;
    dL_dmeans3D[3 * i + 2] = dmz;
    /* ---- A.8 cov3D backward ---- */
    if (in->scales && !in->cov3D_precomp) {
      const float *sc = in->scales + 3 * i;
      const float *q = in->rotations + 4 * i;
      float mod = in->scale_modifier;
      float s[3] = {mod * sc[0], mod * sc[1], mod * sc[2]};
      float r = q[0], x = q[1], y = q[2], z = q[3];
      float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                       {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                       {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
      float G[3][3] = {{dc3[0], 0.5f * dc3[1], 0.5f * dc3[2]},
                       {0.5f * dc3[1], dc3[3], 0.5f * dc3[4]},
                       {0.5f * dc3[2], 0.5f * dc3[4], dc3[5]}};
      float A[3][3], dA[3][3];
      for (int rr = 0; rr < 3; rr++)
        for (int k = 0; k < 3; k++) A[rr][k] = R[rr][k] * s[k];
      for (int rr = 0; rr < 3; rr++)
        for (int k = 0; k < 3; k++)
          dA[rr][k] = 2.f * (G[rr][0] * A[0][k] + G[rr][1] * A[1][k] + G[rr][2] * A[2][k]);
      for (int k = 0; k < 3; k++)
        dL_dscales[3 * i + k] = R[0][k] * dA[0][k] + R[1][k] * dA[1][k] + R[2][k] * dA[2][k];
      float Dm[3][3];
      for (int rr = 0; rr < 3; rr++)
        for (int k = 0; k < 3; k++) Dm[rr][k] = dA[rr][k] * s[k];
      dL_drots[4 * i + 0] = 2 * z * (Dm[1][0] - Dm[0][1]) + 2 * y * (Dm[0][2] - Dm[2][0]) + 2 * x * (Dm[2][1] - Dm[1][2]);
      dL_drots[4 * i + 1] = 2 * y * (Dm[0][1] + Dm[1][0]) + 2 * z * (Dm[0][2] + Dm[2][0]) + 2 * r * (Dm[2][1] - Dm[1][2]) - 4 * x * (Dm[1][1] + Dm[2][2]);
      dL_drots[4 * i + 2] = 2 * x * (Dm[0][1] + Dm[1][0]) + 2 * r * (Dm[0][2] - Dm[2][0]) + 2 * z * (Dm[1][2] + Dm[2][1]) - 4 * y * (Dm[0][0] + Dm[2][2]);
      dL_drots[4 * i + 3] = 2 * r * (Dm[1][0] - Dm[0][1]) + 2 * x * (Dm[0][2] + Dm[2][0]) + 2 * y * (Dm[1][2] + Dm[2][1]) - 4 * z * (Dm[0][0] + Dm[1][1]);
    }
  }
  free(dmean2D); free(dconic); free(dopac); free(dcolor);
  return 0;
}

/* A.2 step 1 only: visibility test used by GaussianRasterizer.markVisible */
GSO_EXPORT void gso_mark_visible(int P, const float *means3D, const float *view, const float *proj,
                                 uint8_t *present) {
  (void)proj;
  for (int i = 0; i < P; i++) {
    const float *p = means3D + 3 * i;
    float pvz = dot3a(view[2], p[0], view[6], p[1], view[10], p[2], view[14]);
    present[i] = (pvz > 0.2f);
  }
}

GSO_EXPORT int gso_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
GSO_EXPORT void gso_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
