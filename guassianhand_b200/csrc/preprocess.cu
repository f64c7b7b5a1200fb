// Per-(view, Gaussian) preprocess: frustum cull, 3D covariance, EWA projection to the 2D conic,
// radius / tile rectangle, SH colour, and the per-tile instance counts the binning starts from.
// Replaces upstream preprocessCUDA<3> -- one thread per
// (view, Gaussian); output is one 64-byte record (4 x STG.128) instead of seven scattered arrays.
#include "ghr_internal.cuh"

namespace ghr {

namespace {

__constant__ float SH_C2c[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                -1.0925484305920792f, 0.5462742152960396f};
__constant__ float SH_C3c[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                -0.5900435899266435f};
constexpr float SH_C0 = 0.28209479177387814f;
constexpr float SH_C1 = 0.4886025119029199f;

// A.2 step 3
__device__ __forceinline__ void cov3d_from_scale_rot(const float *scale, float mod, const float4 q, float *c) {
  float s0 = fmul(mod, scale[0]), s1 = fmul(mod, scale[1]), s2 = fmul(mod, scale[2]);
  float r = q.x, x = q.y, y = q.z, z = q.w;
  float R00 = ffma(-2.f, ffma(y, y, fmul(z, z)), 1.f);
  float R01 = fmul(2.f, ffma(x, y, -fmul(r, z)));
  float R02 = fmul(2.f, ffma(x, z, fmul(r, y)));
  float R10 = fmul(2.f, ffma(x, y, fmul(r, z)));
  float R11 = ffma(-2.f, ffma(x, x, fmul(z, z)), 1.f);
  float R12 = fmul(2.f, ffma(y, z, -fmul(r, x)));
  float R20 = fmul(2.f, ffma(x, z, -fmul(r, y)));
  float R21 = fmul(2.f, ffma(y, z, fmul(r, x)));
  float R22 = ffma(-2.f, ffma(x, x, fmul(y, y)), 1.f);
  float A00 = fmul(R00, s0), A01 = fmul(R01, s1), A02 = fmul(R02, s2);
  float A10 = fmul(R10, s0), A11 = fmul(R11, s1), A12 = fmul(R12, s2);
  float A20 = fmul(R20, s0), A21 = fmul(R21, s1), A22 = fmul(R22, s2);
  c[0] = dot3(A00, A00, A01, A01, A02, A02);
  c[1] = dot3(A00, A10, A01, A11, A02, A12);
  c[2] = dot3(A00, A20, A01, A21, A02, A22);
  c[3] = dot3(A10, A10, A11, A11, A12, A12);
  c[4] = dot3(A10, A20, A11, A21, A12, A22);
  c[5] = dot3(A20, A20, A21, A21, A22, A22);
}

// A.3
__device__ __forceinline__ void sh_to_rgb(int D, const float *sh, float px, float py, float pz,
                                          const float *campos, float *rgb, uint32_t *clamp_bits) {
  float dx = fsub(px, campos[0]), dy = fsub(py, campos[1]), dz = fsub(pz, campos[2]);
  float len = fsqrt(dot3(dx, dx, dy, dy, dz, dz));
  float x = fdiv(dx, len), y = fdiv(dy, len), z = fdiv(dz, len);
  uint32_t bits = 0;
#pragma unroll
  for (int ch = 0; ch < 3; ch++) {
    float r = fmul(SH_C0, sh[ch]);
    if (D > 0) {
      r = ffma(-fmul(SH_C1, y), sh[1 * 3 + ch], r);
      r = ffma(fmul(SH_C1, z), sh[2 * 3 + ch], r);
      r = ffma(-fmul(SH_C1, x), sh[3 * 3 + ch], r);
      if (D > 1) {
        float xx = fmul(x, x), yy = fmul(y, y), zz = fmul(z, z);
        float xy = fmul(x, y), yz = fmul(y, z), xz = fmul(x, z);
        r = ffma(fmul(SH_C2c[0], xy), sh[4 * 3 + ch], r);
        r = ffma(fmul(SH_C2c[1], yz), sh[5 * 3 + ch], r);
        r = ffma(fmul(SH_C2c[2], fsub(ffma(2.0f, zz, -xx), yy)), sh[6 * 3 + ch], r);
        r = ffma(fmul(SH_C2c[3], xz), sh[7 * 3 + ch], r);
        r = ffma(fmul(SH_C2c[4], fsub(xx, yy)), sh[8 * 3 + ch], r);
        if (D > 2) {
          r = ffma(fmul(fmul(SH_C3c[0], y), ffma(3.0f, xx, -yy)), sh[9 * 3 + ch], r);
          r = ffma(fmul(fmul(SH_C3c[1], xy), z), sh[10 * 3 + ch], r);
          r = ffma(fmul(fmul(SH_C3c[2], y), fsub(ffma(4.0f, zz, -xx), yy)), sh[11 * 3 + ch], r);
          r = ffma(fmul(fmul(SH_C3c[3], z), ffma(-3.0f, yy, ffma(2.0f, zz, -fmul(3.0f, xx)))), sh[12 * 3 + ch], r);
          r = ffma(fmul(fmul(SH_C3c[4], x), fsub(ffma(4.0f, zz, -xx), yy)), sh[13 * 3 + ch], r);
          r = ffma(fmul(fmul(SH_C3c[5], z), fsub(xx, yy)), sh[14 * 3 + ch], r);
          r = ffma(fmul(fmul(SH_C3c[6], x), ffma(-3.0f, yy, xx)), sh[15 * 3 + ch], r);
        }
      }
    }
    r = fadd(r, 0.5f);
    if (r < 0.f) bits |= (1u << ch);
    rgb[ch] = fmaxf(r, 0.f);
  }
  *clamp_bits = bits;
}

// A.3 with the coefficient row read as 128-bit loads (rows of M*3 floats that are a multiple of 16
// bytes: M = 4, 16).  Same arithmetic as sh_to_rgb, coefficient by coefficient: channel ch accumulates
// r = C0*sh[0][ch], then r = fma(basis_k, sh[k][ch], r) for k = 1.. in order, with basis_k the very
// expressions of the scalar version -- the rgb bits are identical.
__device__ __forceinline__ void sh_to_rgb_vec(int D, const float4 *__restrict__ sh4, float px, float py, float pz,
                                              const float *campos, float *rgb, uint32_t *clamp_bits) {
  float dx = fsub(px, campos[0]), dy = fsub(py, campos[1]), dz = fsub(pz, campos[2]);
  float len = fsqrt(dot3(dx, dx, dy, dy, dz, dz));
  float x = fdiv(dx, len), y = fdiv(dy, len), z = fdiv(dz, len);
  float basis[16];
  basis[0] = SH_C0;
  {
    float xx = fmul(x, x), yy = fmul(y, y), zz = fmul(z, z);
    float xy = fmul(x, y), yz = fmul(y, z), xz = fmul(x, z);
    basis[1] = -fmul(SH_C1, y);
    basis[2] = fmul(SH_C1, z);
    basis[3] = -fmul(SH_C1, x);
    basis[4] = fmul(SH_C2c[0], xy);
    basis[5] = fmul(SH_C2c[1], yz);
    basis[6] = fmul(SH_C2c[2], fsub(ffma(2.0f, zz, -xx), yy));
    basis[7] = fmul(SH_C2c[3], xz);
    basis[8] = fmul(SH_C2c[4], fsub(xx, yy));
    basis[9] = fmul(fmul(SH_C3c[0], y), ffma(3.0f, xx, -yy));
    basis[10] = fmul(fmul(SH_C3c[1], xy), z);
    basis[11] = fmul(fmul(SH_C3c[2], y), fsub(ffma(4.0f, zz, -xx), yy));
    basis[12] = fmul(fmul(SH_C3c[3], z), ffma(-3.0f, yy, ffma(2.0f, zz, -fmul(3.0f, xx))));
    basis[13] = fmul(fmul(SH_C3c[4], x), fsub(ffma(4.0f, zz, -xx), yy));
    basis[14] = fmul(fmul(SH_C3c[5], z), fsub(xx, yy));
    basis[15] = fmul(fmul(SH_C3c[6], x), ffma(-3.0f, yy, xx));
  }
  const int nf = 3 * (D + 1) * (D + 1);      // floats of the row that are used
  float r[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 12; j++) {
    if (4 * j < nf) {
      const float4 v = sh4[j];
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const int f = 4 * j + c, k = f / 3, ch = f % 3;
        if (f < nf) r[ch] = k == 0 ? fmul(SH_C0, e[c]) : ffma(basis[k], e[c], r[ch]);
      }
    }
  }
  uint32_t bits = 0;
#pragma unroll
  for (int ch = 0; ch < 3; ch++) {
    const float t = fadd(r[ch], 0.5f);
    if (t < 0.f) bits |= (1u << ch);
    rgb[ch] = fmaxf(t, 0.f);
  }
  *clamp_bits = bits;
}

template <bool kVecSH>
__global__ void __launch_bounds__(256)
preprocess_kernel(int P, int V, int H, int W, int M, int D, int gx, int gy, float scale_modifier, uint32_t flags,
                  Cameras cam, Gaussians g, float4 *__restrict__ geom, uint8_t *__restrict__ clamped,
                  int32_t *__restrict__ radii, uint32_t *__restrict__ tile_count, uint32_t *__restrict__ tilemax,
                  uint32_t *__restrict__ misc, int T, int smem_tiles) {
  // s_cnt[t]: instances this block's Gaussians put on tile t of its view (smem_tiles == T), flushed
  // as one RED per touched tile; with more tiles than fit (smem_tiles == 0) every instance is a RED
  extern __shared__ uint32_t s_cnt[];
  const int v = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  for (int k = threadIdx.x; k < smem_tiles; k += blockDim.x) s_cnt[k] = 0;
  // zero the per-tile blend output (grid-stride over all threads of the launch)
  {
    size_t gtid = ((size_t)v * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    size_t gsz = (size_t)gridDim.x * gridDim.y * blockDim.x;
    for (size_t t = gtid; t < (size_t)2 * T * V; t += gsz) tilemax[t] = 0u;
  }
  __syncthreads();

  bool visible = false;
  if (i < P) {
    const float *Vm = cam.view + 16 * v;
    const float *PV = cam.proj + 16 * v;
    float tanfovx = cam.tanfov ? cam.tanfov[2 * v] : cam.tanfovx;
    float tanfovy = cam.tanfov ? cam.tanfov[2 * v + 1] : cam.tanfovy;
    // (12-byte rows: the three scalar loads of a warp use every byte of the sectors they touch; staging the
    // rows through shared memory with 128-bit loads was measured 3 us slower -- the kernel is latency-bound)
    float px = g.means3D[3 * i], py = g.means3D[3 * i + 1], pz = g.means3D[3 * i + 2];
    int radius = 0;
    uint32_t tiles = 0;
    float4 q0 = make_float4(0, 0, 0, 0), q1 = q0, q2 = q0, q3 = q0;
    q1.z = -1.f;   // cull threshold: never contributes
    uint32_t cbits = 0;

    float pvx = dot3a(Vm[0], px, Vm[4], py, Vm[8], pz, Vm[12]);
    float pvy = dot3a(Vm[1], px, Vm[5], py, Vm[9], pz, Vm[13]);
    float pvz = dot3a(Vm[2], px, Vm[6], py, Vm[10], pz, Vm[14]);
    if (pvz > 0.2f) {
      float phx = dot3a(PV[0], px, PV[4], py, PV[8], pz, PV[12]);
      float phy = dot3a(PV[1], px, PV[5], py, PV[9], pz, PV[13]);
      float phw = dot3a(PV[3], px, PV[7], py, PV[11], pz, PV[15]);
      float p_w = fdiv(1.0f, fadd(phw, 0.0000001f));
      float ppx = fmul(phx, p_w), ppy = fmul(phy, p_w);
      float c3[6];
      if (g.cov3D_precomp) {
#pragma unroll
        for (int k = 0; k < 6; k++) c3[k] = g.cov3D_precomp[6 * (size_t)i + k];
      } else {
        float sc[3] = {g.scales[3 * i], g.scales[3 * i + 1], g.scales[3 * i + 2]};
        float4 q = reinterpret_cast<const float4 *>(g.rotations)[i];
        cov3d_from_scale_rot(sc, scale_modifier, q, c3);
      }
      // A.2 step 4
      float fx = fdiv((float)W, fmul(2.0f, tanfovx));
      float fy = fdiv((float)H, fmul(2.0f, tanfovy));
      float limx = fmul(1.3f, tanfovx), limy = fmul(1.3f, tanfovy);
      float txtz = fdiv(pvx, pvz), tytz = fdiv(pvy, pvz);
      float tx = fmul(fminf(limx, fmaxf(-limx, txtz)), pvz);
      float ty = fmul(fminf(limy, fmaxf(-limy, tytz)), pvz);
      float tz = pvz;
      float J00 = fdiv(fx, tz);
      float J02 = fdiv(-fmul(fx, tx), fmul(tz, tz));
      float J11 = fdiv(fy, tz);
      float J12 = fdiv(-fmul(fy, ty), fmul(tz, tz));
      float T0[3], T1[3];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        float W0 = Vm[4 * c + 0], W1 = Vm[4 * c + 1], W2 = Vm[4 * c + 2];
        T0[c] = dot2(J00, W0, J02, W2);
        T1[c] = dot2(J11, W1, J12, W2);
      }
      float u0[3], u1[3];
      const float S[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
#pragma unroll
      for (int r = 0; r < 3; r++) {
        u0[r] = dot3(S[r][0], T0[0], S[r][1], T0[1], S[r][2], T0[2]);
        u1[r] = dot3(S[r][0], T1[0], S[r][1], T1[1], S[r][2], T1[2]);
      }
      float a = fadd(dot3(T0[0], u0[0], T0[1], u0[1], T0[2], u0[2]), 0.3f);
      float b = dot3(T0[0], u1[0], T0[1], u1[1], T0[2], u1[2]);
      float c = fadd(dot3(T1[0], u1[0], T1[1], u1[1], T1[2], u1[2]), 0.3f);
      float det = ffma(a, c, -fmul(b, b));
      if (det != 0.0f) {
        float det_inv = fdiv(1.f, det);
        float conx = fmul(c, det_inv), cony = fmul(-b, det_inv), conz = fmul(a, det_inv);
        float mid = fmul(0.5f, fadd(a, c));
        float sq = fsqrt(fmaxf(0.1f, ffma(mid, mid, -det)));
        float l1 = fadd(mid, sq), l2 = fsub(mid, sq);
        int rad = (int)ceilf(fmul(3.f, fsqrt(fmaxf(l1, l2))));
        float pix_x = fmul(ffma(fadd(ppx, 1.0f), (float)W, -1.0f), 0.5f);
        float pix_y = fmul(ffma(fadd(ppy, 1.0f), (float)H, -1.0f), 0.5f);
        float rf = (float)rad;
        // x / 16 == x * 0.0625 bit for bit (power of two)
        int minx = min(gx, max(0, (int)fmul(fsub(pix_x, rf), 0.0625f)));
        int miny = min(gy, max(0, (int)fmul(fsub(pix_y, rf), 0.0625f)));
        int maxx = min(gx, max(0, (int)fmul(fsub(fadd(fadd(pix_x, rf), 16.0f), 1.0f), 0.0625f)));
        int maxy = min(gy, max(0, (int)fmul(fsub(fadd(fadd(pix_y, rf), 16.0f), 1.0f), 0.0625f)));
        int tt = (maxx - minx) * (maxy - miny);
        if (tt != 0) {
          float rgb[3];
          if (g.colors_precomp) {
            rgb[0] = g.colors_precomp[3 * i];
            rgb[1] = g.colors_precomp[3 * i + 1];
            rgb[2] = g.colors_precomp[3 * i + 2];
          } else if (kVecSH) {
            sh_to_rgb_vec(D, reinterpret_cast<const float4 *>(g.shs + (size_t)i * M * 3), px, py, pz,
                          cam.campos + 3 * v, rgb, &cbits);
          } else {
            sh_to_rgb(D, g.shs + (size_t)i * M * 3, px, py, pz, cam.campos + 3 * v, rgb, &cbits);
          }
          radius = rad;
          tiles = (uint32_t)tt;
          visible = true;
          uint32_t *cnt = smem_tiles ? s_cnt : tile_count + (size_t)v * T;
          for (int y = miny; y < maxy; y++)
            for (int x = minx; x < maxx; x++) atomicAdd(&cnt[y * gx + x], 1u);
          // Cull threshold for the blend kernels (never part of the canonical arithmetic):
          // alpha = o*exp(power) >= 1/255  =>  d^T Q d <= 2 ln(255 o) =: thr.  A warp skips the
          // instance when the minimum of d^T Q d over its pixel block exceeds thr (with margin).
          // thr < 0: can never contribute; thr = 3e38: ill-conditioned conic, never cull.
          const float opac = g.opacities[i];
          float thr = -1.f;
          if (opac >= 0.0039f) {
            float tau = fmaxf(logf(255.f * opac), 0.f) + 1e-3f;
            float detq = conx * conz - cony * cony;
            thr = (conx > 0.f && conz > 0.f && detq > 1e-3f * conx * conz) ? 2.f * tau * 1.01f + 0.02f : 3e38f;
          }
          q0 = make_float4(pix_x, pix_y, conx, cony);
          q1 = make_float4(conz, opac, thr, 0.f);
          q2 = make_float4(rgb[0], rgb[1], rgb[2], pvz);
          q3 = make_float4(__int_as_float(rad), __uint_as_float(tiles), 0.f, 0.f);
        }
      }
    } else if (flags & GHR_FLAG_PREFILTERED) {
      // upstream prints "Point is filtered although prefiltered is set" and __trap()s here, which takes the CUDA
      // context down; the point is culled and the violation reported in GhrStatus.overflow (GHR_STATUS_PREFILTER;
      // the tile scan assembles the status from the zero-initialised misc words)
      atomicOr(&misc[kMiscPrefilter], 1u);
    }
    size_t e = (size_t)v * P + i;
    geom[4 * e + 0] = q0;
    geom[4 * e + 1] = q1;
    geom[4 * e + 2] = q2;
    geom[4 * e + 3] = q3;
    if (clamped) clamped[e] = (uint8_t)cbits;
    radii[e] = radius;
  }
  const uint32_t nvis = __popc(__ballot_sync(0xFFFFFFFFu, visible));
  if ((threadIdx.x & 31) == 0 && nvis) atomicAdd(&misc[kMiscVisible], nvis);
  if (smem_tiles) {
    __syncthreads();
    for (int k = threadIdx.x; k < smem_tiles; k += blockDim.x) {
      const uint32_t c = s_cnt[k];
      if (c) atomicAdd(&tile_count[(size_t)v * T + k], c);
    }
  }
}

// Geometry reuse (GhrForwardArgs.reuse_state): the projected record of every Gaussian is copied from the
// earlier call's state and only its colour is recomputed (colors_precomp, or the SH evaluation in the very
// order of preprocess_kernel); radii are re-exported.  One thread per Gaussian, one view.
template <bool kVecSH>
__global__ void __launch_bounds__(256)
recolor_geom_kernel(int P, int M, int D, Cameras cam, Gaussians g, const float4 *__restrict__ old_geom,
                    float4 *__restrict__ geom, uint8_t *__restrict__ clamped, int32_t *__restrict__ radii) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float4 q0 = old_geom[4 * (size_t)i], q1 = old_geom[4 * (size_t)i + 1], q3 = old_geom[4 * (size_t)i + 3];
  float4 q2 = old_geom[4 * (size_t)i + 2];
  uint32_t cbits = 0;
  if (__float_as_uint(q3.y)) {               // visible: tiles_touched > 0
    float rgb[3];
    if (g.colors_precomp) {
      rgb[0] = g.colors_precomp[3 * i];
      rgb[1] = g.colors_precomp[3 * i + 1];
      rgb[2] = g.colors_precomp[3 * i + 2];
    } else {
      const float px = g.means3D[3 * i], py = g.means3D[3 * i + 1], pz = g.means3D[3 * i + 2];
      if (kVecSH) sh_to_rgb_vec(D, reinterpret_cast<const float4 *>(g.shs + (size_t)i * M * 3), px, py, pz, cam.campos, rgb, &cbits);
      else sh_to_rgb(D, g.shs + (size_t)i * M * 3, px, py, pz, cam.campos, rgb, &cbits);
    }
    q2.x = rgb[0];
    q2.y = rgb[1];
    q2.z = rgb[2];
  }
  geom[4 * (size_t)i] = q0;
  geom[4 * (size_t)i + 1] = q1;
  geom[4 * (size_t)i + 2] = q2;
  geom[4 * (size_t)i + 3] = q3;
  if (clamped) clamped[i] = (uint8_t)cbits;
  radii[i] = __float_as_int(q3.x);
}

__global__ void mark_visible_kernel(int P, const float *__restrict__ means3D, const float *__restrict__ Vm,
                                    uint8_t *__restrict__ present) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  float pvz = dot3a(Vm[2], means3D[3 * i], Vm[6], means3D[3 * i + 1], Vm[10], means3D[3 * i + 2], Vm[14]);
  present[i] = pvz > 0.2f;
}

// Cameras of V views from (w2c, K) on the device: what the reference builds per view on the host with two
// implicit device->host syncs (/root/reference/tgs/models/renderer_one_shot.py:61-112 Camera.from_w2c +
// getProjectionMatrix_refine + intrinsic_to_fov, :278-279 math.tan of a CUDA scalar).  One thread per view.
//   viewmatrix = w2c^T;  projmatrix = viewmatrix @ P^T with P from the intrinsics (off-centre cx, cy, skew);
//   campos = inverse(viewmatrix)[3, :3];  tanfov = tan(0.5 * 2 atan2(size, 2 f)) evaluated like the reference
//   (fp32 atan2, fp32 halving, double tan, rounded to fp32).
__global__ void cameras_from_w2c_kernel(int V, const float *__restrict__ w2c, const float *__restrict__ K, int H, int W,
                                        float znear, float zfar, float *__restrict__ view, float *__restrict__ proj,
                                        float *__restrict__ campos, float *__restrict__ tanfov) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const float *m = w2c + 16 * v, *k = K + 9 * v;
  float vw[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) vw[i][j] = m[4 * j + i];            // transpose
  const float fx = k[0], sk = k[1], cx = k[2], fy = k[4], cy = k[5];
  float P[4][4] = {};
  P[0][0] = fdiv(fmul(2.f, fx), (float)W);
  P[0][1] = fdiv(fmul(2.f, sk), (float)W);
  P[0][2] = fadd(-1.f, fmul(2.f, fdiv(cx, (float)W)));
  P[1][1] = fdiv(fmul(2.f, fy), (float)H);
  P[1][2] = fadd(-1.f, fmul(2.f, fdiv(cy, (float)H)));
  P[2][2] = fdiv(fadd(zfar, znear), fsub(zfar, znear));
  P[2][3] = fdiv(fmul(fmul(-2.f, zfar), znear), fsub(zfar, znear));
  P[3][2] = 1.f;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float acc = fmul(vw[i][0], P[j][0]);                            // (view @ P^T)[i][j] = sum_k view[i][k] P[j][k]
#pragma unroll
      for (int q = 1; q < 4; q++) acc = ffma(vw[i][q], P[j][q], acc);
      proj[16 * v + 4 * i + j] = acc;
      view[16 * v + 4 * i + j] = vw[i][j];
    }
  // camera centre: inverse(view)[3, :3] = -(A^-1 t) for w2c = [A t; 0 0 0 1], in double (adjugate / determinant)
  const double a00 = m[0], a01 = m[1], a02 = m[2], a10 = m[4], a11 = m[5], a12 = m[6], a20 = m[8], a21 = m[9], a22 = m[10];
  const double t0 = m[3], t1 = m[7], t2 = m[11];
  const double c00 = a11 * a22 - a12 * a21, c01 = a02 * a21 - a01 * a22, c02 = a01 * a12 - a02 * a11;
  const double c10 = a12 * a20 - a10 * a22, c11 = a00 * a22 - a02 * a20, c12 = a02 * a10 - a00 * a12;
  const double c20 = a10 * a21 - a11 * a20, c21 = a01 * a20 - a00 * a21, c22 = a00 * a11 - a01 * a10;
  const double det = a00 * c00 + a01 * c10 + a02 * c20, id = 1.0 / det;
  campos[3 * v + 0] = (float)(-(c00 * t0 + c01 * t1 + c02 * t2) * id);
  campos[3 * v + 1] = (float)(-(c10 * t0 + c11 * t1 + c12 * t2) * id);
  campos[3 * v + 2] = (float)(-(c20 * t0 + c21 * t1 + c22 * t2) * id);
  const float fovx = fmul(2.f, atan2f((float)W, fmul(2.f, fx))), fovy = fmul(2.f, atan2f((float)H, fmul(2.f, fy)));
  tanfov[2 * v + 0] = (float)tan((double)fmul(fovx, 0.5f));
  tanfov[2 * v + 1] = (float)tan((double)fmul(fovy, 0.5f));
}

}  // namespace

cudaError_t launch_cameras_from_w2c(int V, const float *w2c, const float *K, int H, int W, float znear, float zfar,
                                    float *view, float *proj, float *campos, float *tanfov, cudaStream_t s) {
  if (V <= 0) return cudaSuccess;
  cameras_from_w2c_kernel<<<(V + 63) / 64, 64, 0, s>>>(V, w2c, K, H, W, znear, zfar, view, proj, campos, tanfov);
  return cudaGetLastError();
}

cudaError_t launch_preprocess(const GhrDims &d, const Layout &L, const Cameras &cam, const Gaussians &g,
                              float scale_modifier, uint32_t flags, char *state, char *temp, int32_t *radii,
                              cudaStream_t s) {
  int nb = (d.P + 255) / 256;
  if (nb == 0) nb = 1;
  dim3 grid(nb, d.V), block(256);
  const int smem_tiles = L.T <= kMaxSmemTiles ? L.T : 0;
  const size_t smem = (size_t)smem_tiles * 4;
  // SH rows that are a multiple of 16 bytes (M = 4, 16) on a 16-byte aligned base take the 128-bit path
  const bool vec_sh = g.shs && !g.colors_precomp && (d.M * 3) % 4 == 0 && ((uintptr_t)g.shs & 15) == 0;
  auto launch = [&](auto kern) -> cudaError_t {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    kern<<<grid, block, smem, s>>>(
        d.P, d.V, d.H, d.W, d.M, d.sh_degree, L.gx, L.gy, scale_modifier, flags, cam, g,
        (float4 *)(state + L.pub.off_geom), d.M > 0 ? (uint8_t *)(state + L.pub.off_clamped) : nullptr, radii,
        (uint32_t *)(temp + L.t_tile_count), (uint32_t *)(state + L.pub.off_tilemax),
        (uint32_t *)(temp + L.t_misc), L.T, smem_tiles);
    return cudaSuccess;
  };
  cudaError_t e = vec_sh ? launch(preprocess_kernel<true>) : launch(preprocess_kernel<false>);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

cudaError_t launch_recolor_geom(const GhrDims &d, const Layout &L, const Layout &Lold, const Cameras &cam,
                                const Gaussians &g, const char *old_state, char *state, int32_t *radii, cudaStream_t s) {
  if (d.P == 0) return cudaSuccess;
  const int nb = (d.P + 255) / 256;
  const bool vec_sh = g.shs && !g.colors_precomp && (d.M * 3) % 4 == 0 && ((uintptr_t)g.shs & 15) == 0;
  uint8_t *cl = d.M > 0 ? (uint8_t *)(state + L.pub.off_clamped) : nullptr;
  const float4 *og = (const float4 *)(old_state + Lold.pub.off_geom);
  if (vec_sh) recolor_geom_kernel<true><<<nb, 256, 0, s>>>(d.P, d.M, d.sh_degree, cam, g, og, (float4 *)(state + L.pub.off_geom), cl, radii);
  else recolor_geom_kernel<false><<<nb, 256, 0, s>>>(d.P, d.M, d.sh_degree, cam, g, og, (float4 *)(state + L.pub.off_geom), cl, radii);
  return cudaGetLastError();
}

cudaError_t launch_mark_visible(int P, const float *means3D, const float *view, uint8_t *present,
                                cudaStream_t s) {
  if (P <= 0) return cudaSuccess;
  mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, view, present);
  return cudaGetLastError();
}

}  // namespace ghr
