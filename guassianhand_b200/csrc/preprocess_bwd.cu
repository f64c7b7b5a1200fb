// Backward of the per-Gaussian preprocess.  Replaces upstream computeCov2DCUDA + preprocessCUDA<3>
// backward (SURVEY.md §2a K8,K9; A.7, A.8, A.3) with ONE kernel: one thread per Gaussian walks the
// V views, converts the blend-stage moment accumulators into dL/dmean2D, dL/dconic, dL/dopacity,
// pushes them through the EWA projection, the perspective divide, the SH evaluation and the
// scale/quaternion covariance, and sums over views in registers -- so the multi-view gradient is
// produced deterministically with no atomics and is written (or accumulated) exactly once.
#include "ghr_internal.cuh"

namespace ghr {

namespace {

__constant__ float SHB_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                -1.0925484305920792f, 0.5462742152960396f};
__constant__ float SHB_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                -0.5900435899266435f};
constexpr float SH_C0 = 0.28209479177387814f;
constexpr float SH_C1 = 0.4886025119029199f;

// A.2 step 3 in the canonical order (same bits as preprocess.cu), also returning R and s
__device__ __forceinline__ void cov3d_of(const float *sc, float mod, float4 q, float *c3, float (*R)[3], float *s) {
  s[0] = fmul(mod, sc[0]); s[1] = fmul(mod, sc[1]); s[2] = fmul(mod, sc[2]);
  float r = q.x, x = q.y, y = q.z, z = q.w;
  R[0][0] = ffma(-2.f, ffma(y, y, fmul(z, z)), 1.f);
  R[0][1] = fmul(2.f, ffma(x, y, -fmul(r, z)));
  R[0][2] = fmul(2.f, ffma(x, z, fmul(r, y)));
  R[1][0] = fmul(2.f, ffma(x, y, fmul(r, z)));
  R[1][1] = ffma(-2.f, ffma(x, x, fmul(z, z)), 1.f);
  R[1][2] = fmul(2.f, ffma(y, z, -fmul(r, x)));
  R[2][0] = fmul(2.f, ffma(x, z, -fmul(r, y)));
  R[2][1] = fmul(2.f, ffma(y, z, fmul(r, x)));
  R[2][2] = ffma(-2.f, ffma(x, x, fmul(y, y)), 1.f);
  float A[3][3];
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int k = 0; k < 3; k++) A[a][k] = fmul(R[a][k], s[k]);
  c3[0] = dot3(A[0][0], A[0][0], A[0][1], A[0][1], A[0][2], A[0][2]);
  c3[1] = dot3(A[0][0], A[1][0], A[0][1], A[1][1], A[0][2], A[1][2]);
  c3[2] = dot3(A[0][0], A[2][0], A[0][1], A[2][1], A[0][2], A[2][2]);
  c3[3] = dot3(A[1][0], A[1][0], A[1][1], A[1][1], A[1][2], A[1][2]);
  c3[4] = dot3(A[1][0], A[2][0], A[1][1], A[2][1], A[1][2], A[2][2]);
  c3[5] = dot3(A[2][0], A[2][0], A[2][1], A[2][1], A[2][2], A[2][2]);
}

#ifndef GHR_PBWD_THREADS
#define GHR_PBWD_THREADS 128
#endif

// kVecSH: the SH coefficient row and its gradient row are multiples of 16 bytes on 16-byte aligned bases
// (M = 4, 16): both are streamed with 128-bit accesses, 12 instead of 48 per view for degree 3.
template <bool HAS_SH, bool kVecSH>
__global__ void __launch_bounds__(GHR_PBWD_THREADS)
preprocess_backward_kernel(int P, int V, int H, int W, int M, int D, float scale_modifier, Cameras cam,
                           Gaussians g, const float4 *__restrict__ geom, const uint8_t *__restrict__ clamped,
                           const float *__restrict__ acc, GradOut go) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float px = g.means3D[3 * i], py = g.means3D[3 * i + 1], pz = g.means3D[3 * i + 2];
  const bool from_sr = g.cov3D_precomp == nullptr;
  float c3[6], R[3][3], s[3];
  if (from_sr) {
    float sc[3] = {g.scales[3 * i], g.scales[3 * i + 1], g.scales[3 * i + 2]};
    cov3d_of(sc, scale_modifier, reinterpret_cast<const float4 *>(g.rotations)[i], c3, R, s);
  } else {
#pragma unroll
    for (int k = 0; k < 6; k++) c3[k] = g.cov3D_precomp[6 * (size_t)i + k];
  }
  const float opac = g.opacities[i];

  float dmx = 0.f, dmy = 0.f, dmz = 0.f, dop = 0.f;
  float dcol[3] = {0.f, 0.f, 0.f};
  float dc3[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float *dsh_out = HAS_SH && go.dsh ? go.dsh + (size_t)i * M * 3 : nullptr;
  // SH gradients: the 128-bit variant sums a Gaussian's gradient row over the views in shared memory (48 floats per
  // thread, [coefficient][thread]: conflict-free) and writes it once -- a read-modify-write of the 192-byte output
  // row per view was 3/4 of this kernel's traffic at 1M Gaussians; the scalar variant accumulates in the output row
  constexpr int kShAcc = kVecSH ? 48 : 1;
  __shared__ float s_dsh[kShAcc][GHR_PBWD_THREADS];
  if (kVecSH && dsh_out) {
#pragma unroll
    for (int f = 0; f < 48; f++) s_dsh[f][threadIdx.x] = 0.f;
  }
  bool sh_first = true;

  // The per-view loads are software-pipelined: one thread walks its Gaussian's V views serially (the view sum
  // stays in registers: deterministic, no atomics) and the kernel runs at ~20 % occupancy, so a view's loads must
  // be in flight while the previous view is computed.  The radius (which gates the other loads: a culled
  // (view, Gaussian) reads one float4) runs two views ahead, the record and the accumulator row one view ahead.
  struct ViewLoads {
    float4 q0, q1, a0, a1;
    float m02;
  };
  auto load_view = [&](int v_, ViewLoads &L_) {
    const size_t e_ = (size_t)v_ * P + i;
    L_.q0 = geom[4 * e_ + 0];
    L_.q1 = geom[4 * e_ + 1];
    const float *a_ = acc + e_ * kAccStride;
    L_.a0 = *reinterpret_cast<const float4 *>(a_);
    L_.a1 = *reinterpret_cast<const float4 *>(a_ + 4);
    L_.m02 = a_[8];
  };
  // (SH variants: the per-view SH arithmetic dominates and the extra live registers cost occupancy -- 167
  // registers for the 128-bit variant -- so they load at the point of use)
  constexpr bool kPipe = !HAS_SH;
  int rad_cur = 0, rad_next = 0;
  ViewLoads cur_l, next_l;
  if (kPipe) {
    rad_cur = __float_as_int(geom[4 * (size_t)i + 3].x);
    rad_next = V > 1 ? __float_as_int(geom[4 * ((size_t)P + i) + 3].x) : 0;
    if (rad_cur > 0) load_view(0, cur_l);
  }

  for (int v = 0; v < V; v++) {
    const size_t e = (size_t)v * P + i;
    int radius;
    ViewLoads vl;
    if (kPipe) {
      radius = rad_cur;
      vl = cur_l;
      // issue the loads of the views ahead before this view's arithmetic
      const int rad_next2 = v + 2 < V ? __float_as_int(geom[4 * ((size_t)(v + 2) * P + i) + 3].x) : 0;
      if (rad_next > 0) load_view(v + 1, next_l);
      rad_cur = rad_next;
      rad_next = rad_next2;
      cur_l = next_l;
    } else {
      radius = __float_as_int(geom[4 * e + 3].x);
      if (radius > 0) load_view(v, vl);
    }
    float g2x = 0.f, g2y = 0.f, gA = 0.f, gB = 0.f, gC = 0.f;
    if (radius > 0) {
      const float4 q0 = vl.q0;
      const float4 q1 = vl.q1;
      const float4 a0 = vl.a0;
      const float4 a1 = vl.a1;
      const float m02 = vl.m02;
      const float dcr = a0.x, dcg = a0.y, dcb = a0.z, m00 = a0.w, m10 = a1.x, m01 = a1.y, m20 = a1.z, m11 = a1.w;
      const float cA = q0.z, cB = q0.w, cC = q1.x;
      // A.6 tail, factored per Gaussian: moments of w = G * dL/dalpha
      g2x = -0.5f * (float)W * opac * (cA * m10 + cB * m01);
      g2y = -0.5f * (float)H * opac * (cC * m01 + cB * m10);
      gA = -0.5f * opac * m20;
      gB = -0.5f * opac * m11;
      gC = -0.5f * opac * m02;
      dop += m00;

      const float *Vm = cam.view + 16 * v;
      const float *PV = cam.proj + 16 * v;
      const float tanfovx = cam.tanfov ? cam.tanfov[2 * v] : cam.tanfovx;
      const float tanfovy = cam.tanfov ? cam.tanfov[2 * v + 1] : cam.tanfovy;
      // ---- A.7 ----
      // recompute the forward quantities in the canonical order (bit-identical to preprocess.cu):
      // denom = a*c - b*b cancels catastrophically for anisotropic splats, so its rounding must
      // be the forward's, not whatever contraction the compiler picks here
      const float tvx = dot3a(Vm[0], px, Vm[4], py, Vm[8], pz, Vm[12]);
      const float tvy = dot3a(Vm[1], px, Vm[5], py, Vm[9], pz, Vm[13]);
      const float tvz = dot3a(Vm[2], px, Vm[6], py, Vm[10], pz, Vm[14]);
      const float fx = fdiv((float)W, fmul(2.0f, tanfovx)), fy = fdiv((float)H, fmul(2.0f, tanfovy));
      const float limx = fmul(1.3f, tanfovx), limy = fmul(1.3f, tanfovy);
      const float txtz = fdiv(tvx, tvz), tytz = fdiv(tvy, tvz);
      const float tx = fmul(fminf(limx, fmaxf(-limx, txtz)), tvz);
      const float ty = fmul(fminf(limy, fmaxf(-limy, tytz)), tvz);
      const float gxm = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
      const float gym = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
      const float itz = 1.f / tvz, itz2 = itz * itz, itz3 = itz2 * itz;
      const float J00 = fdiv(fx, tvz), J02 = fdiv(-fmul(fx, tx), fmul(tvz, tvz));
      const float J11 = fdiv(fy, tvz), J12 = fdiv(-fmul(fy, ty), fmul(tvz, tvz));
      float T0[3], T1[3];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        T0[c] = dot2(J00, Vm[4 * c + 0], J02, Vm[4 * c + 2]);
        T1[c] = dot2(J11, Vm[4 * c + 1], J12, Vm[4 * c + 2]);
      }
      const float S[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
      float u0[3], u1[3];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        u0[r] = dot3(S[r][0], T0[0], S[r][1], T0[1], S[r][2], T0[2]);
        u1[r] = dot3(S[r][0], T1[0], S[r][1], T1[1], S[r][2], T1[2]);
      }
      const float ca = fadd(dot3(T0[0], u0[0], T0[1], u0[1], T0[2], u0[2]), 0.3f);
      const float cb = dot3(T0[0], u1[0], T0[1], u1[1], T0[2], u1[2]);
      const float cc = fadd(dot3(T1[0], u1[0], T1[1], u1[1], T1[2], u1[2]), 0.3f);
      const float denom = ffma(ca, cc, -fmul(cb, cb));
      const float d2i = 1.0f / (denom * denom + 0.0000001f);
      if (d2i != 0.f) {
        const float dL_da = d2i * (-cc * cc * gA + 2.f * cb * cc * gB + (denom - ca * cc) * gC);
        const float dL_dc = d2i * (-ca * ca * gC + 2.f * ca * cb * gB + (denom - ca * cc) * gA);
        const float dL_db = d2i * 2.f * (cb * cc * gA - (denom + 2.f * cb * cb) * gB + ca * cb * gC);
        dc3[0] += T0[0] * T0[0] * dL_da + T0[0] * T1[0] * dL_db + T1[0] * T1[0] * dL_dc;
        dc3[3] += T0[1] * T0[1] * dL_da + T0[1] * T1[1] * dL_db + T1[1] * T1[1] * dL_dc;
        dc3[5] += T0[2] * T0[2] * dL_da + T0[2] * T1[2] * dL_db + T1[2] * T1[2] * dL_dc;
        dc3[1] += 2.f * T0[0] * T0[1] * dL_da + (T0[0] * T1[1] + T0[1] * T1[0]) * dL_db + 2.f * T1[0] * T1[1] * dL_dc;
        dc3[2] += 2.f * T0[0] * T0[2] * dL_da + (T0[0] * T1[2] + T0[2] * T1[0]) * dL_db + 2.f * T1[0] * T1[2] * dL_dc;
        dc3[4] += 2.f * T0[2] * T0[1] * dL_da + (T0[1] * T1[2] + T0[2] * T1[1]) * dL_db + 2.f * T1[1] * T1[2] * dL_dc;
        float dJ00 = 0.f, dJ02 = 0.f, dJ11 = 0.f, dJ12 = 0.f;
#pragma unroll
        for (int r = 0; r < 3; r++) {
          const float dT0 = 2.f * u0[r] * dL_da + u1[r] * dL_db;
          const float dT1 = 2.f * u1[r] * dL_dc + u0[r] * dL_db;
          dJ00 += Vm[4 * r + 0] * dT0;
          dJ02 += Vm[4 * r + 2] * dT0;
          dJ11 += Vm[4 * r + 1] * dT1;
          dJ12 += Vm[4 * r + 2] * dT1;
        }
        const float dtx = gxm * -fx * itz2 * dJ02;
        const float dty = gym * -fy * itz2 * dJ12;
        const float dtz = -fx * itz2 * dJ00 - fy * itz2 * dJ11 + (2.f * fx * tx) * itz3 * dJ02 +
                          (2.f * fy * ty) * itz3 * dJ12;
        dmx += Vm[0] * dtx + Vm[1] * dty + Vm[2] * dtz;
        dmy += Vm[4] * dtx + Vm[5] * dty + Vm[6] * dtz;
        dmz += Vm[8] * dtx + Vm[9] * dty + Vm[10] * dtz;
      }
      // ---- A.8 projection ----
      const float phw = PV[3] * px + PV[7] * py + PV[11] * pz + PV[15];
      const float m_w = 1.0f / (phw + 0.0000001f);
      const float mul1 = (PV[0] * px + PV[4] * py + PV[8] * pz + PV[12]) * m_w * m_w;
      const float mul2 = (PV[1] * px + PV[5] * py + PV[9] * pz + PV[13]) * m_w * m_w;
      dmx += (PV[0] * m_w - PV[3] * mul1) * g2x + (PV[1] * m_w - PV[3] * mul2) * g2y;
      dmy += (PV[4] * m_w - PV[7] * mul1) * g2x + (PV[5] * m_w - PV[7] * mul2) * g2y;
      dmz += (PV[8] * m_w - PV[11] * mul1) * g2x + (PV[9] * m_w - PV[11] * mul2) * g2y;

      if (HAS_SH) {
        // ---- A.3 backward ----
        const float *sh = g.shs + (size_t)i * M * 3;
        const uint8_t cl = clamped[e];
        const float dRGB[3] = {(cl & 1) ? 0.f : dcr, (cl & 2) ? 0.f : dcg, (cl & 4) ? 0.f : dcb};
        const float *cp = cam.campos + 3 * v;
        const float ox = px - cp[0], oy = py - cp[1], oz = pz - cp[2];
        const float sum2 = ox * ox + oy * oy + oz * oz;
        const float ilen = 1.0f / sqrtf(sum2);
        const float x = ox * ilen, y = oy * ilen, z = oz * ilen;
        float ddx = 0.f, ddy = 0.f, ddz = 0.f;
        float basis[16];
        basis[0] = SH_C0;
        if (D > 0) {
          basis[1] = -SH_C1 * y; basis[2] = SH_C1 * z; basis[3] = -SH_C1 * x;
          if (D > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            basis[4] = SHB_C2[0] * xy; basis[5] = SHB_C2[1] * yz; basis[6] = SHB_C2[2] * (2.f * zz - xx - yy);
            basis[7] = SHB_C2[3] * xz; basis[8] = SHB_C2[4] * (xx - yy);
            if (D > 2) {
              basis[9] = SHB_C3[0] * y * (3.f * xx - yy); basis[10] = SHB_C3[1] * xy * z;
              basis[11] = SHB_C3[2] * y * (4.f * zz - xx - yy); basis[12] = SHB_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
              basis[13] = SHB_C3[4] * x * (4.f * zz - xx - yy); basis[14] = SHB_C3[5] * z * (xx - yy);
              basis[15] = SHB_C3[6] * x * (xx - 3.f * yy);
            }
          }
        }
        const int nb = (D + 1) * (D + 1);
        if (kVecSH) {
          // t_k = sum_ch sh[k][ch] * dL/dRGB[ch]: the view-direction gradient needs only these 16 sums
          float tk[16];
#pragma unroll
          for (int k = 0; k < 16; k++) tk[k] = 0.f;
          const float4 *sh4 = reinterpret_cast<const float4 *>(sh);
          const int nf = 3 * nb;
#pragma unroll
          for (int j = 0; j < 12; j++) {
            if (4 * j < nf) {
              const float4 vv = sh4[j];
              const float e[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
              for (int c = 0; c < 4; c++) {
                const int f = 4 * j + c, k = f / 3, ch = f % 3;
                if (f < nf) {
                  if (dsh_out) s_dsh[kVecSH ? f : 0][threadIdx.x] += basis[k] * dRGB[ch];
                  tk[k] += e[c] * dRGB[ch];
                }
              }
            }
          }
          if (D > 0) {
            ddx = -SH_C1 * tk[3]; ddy = -SH_C1 * tk[1]; ddz = SH_C1 * tk[2];
            if (D > 1) {
              const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
              ddx += SHB_C2[0] * y * tk[4] + SHB_C2[2] * 2.f * -x * tk[6] + SHB_C2[3] * z * tk[7] + SHB_C2[4] * 2.f * x * tk[8];
              ddy += SHB_C2[0] * x * tk[4] + SHB_C2[1] * z * tk[5] + SHB_C2[2] * 2.f * -y * tk[6] + SHB_C2[4] * 2.f * -y * tk[8];
              ddz += SHB_C2[1] * y * tk[5] + SHB_C2[2] * 4.f * z * tk[6] + SHB_C2[3] * x * tk[7];
              if (D > 2) {
                ddx += SHB_C3[0] * tk[9] * 6.f * xy + SHB_C3[1] * tk[10] * yz + SHB_C3[2] * tk[11] * -2.f * xy +
                       SHB_C3[3] * tk[12] * -6.f * xz + SHB_C3[4] * tk[13] * (-3.f * xx + 4.f * zz - yy) +
                       SHB_C3[5] * tk[14] * 2.f * xz + SHB_C3[6] * tk[15] * 3.f * (xx - yy);
                ddy += SHB_C3[0] * tk[9] * 3.f * (xx - yy) + SHB_C3[1] * tk[10] * xz +
                       SHB_C3[2] * tk[11] * (-3.f * yy + 4.f * zz - xx) + SHB_C3[3] * tk[12] * -6.f * yz +
                       SHB_C3[4] * tk[13] * -2.f * xy + SHB_C3[5] * tk[14] * -2.f * yz + SHB_C3[6] * tk[15] * -6.f * xy;
                ddz += SHB_C3[1] * tk[10] * xy + SHB_C3[2] * tk[11] * 8.f * yz +
                       SHB_C3[3] * tk[12] * 3.f * (2.f * zz - xx - yy) + SHB_C3[4] * tk[13] * 8.f * xz +
                       SHB_C3[5] * tk[14] * (xx - yy);
              }
            }
          }
        } else {
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
          const float gch = dRGB[ch];
          if (dsh_out) {
            for (int k = 0; k < nb; k++) {
              float val = basis[k] * gch;
              if (sh_first && !go.accumulate) dsh_out[k * 3 + ch] = val;
              else dsh_out[k * 3 + ch] += val;
            }
          }
#define SHV(k) sh[(k) * 3 + ch]
          float rx = 0.f, ry = 0.f, rz = 0.f;
          if (D > 0) {
            rx = -SH_C1 * SHV(3); ry = -SH_C1 * SHV(1); rz = SH_C1 * SHV(2);
            if (D > 1) {
              const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
              rx += SHB_C2[0] * y * SHV(4) + SHB_C2[2] * 2.f * -x * SHV(6) + SHB_C2[3] * z * SHV(7) + SHB_C2[4] * 2.f * x * SHV(8);
              ry += SHB_C2[0] * x * SHV(4) + SHB_C2[1] * z * SHV(5) + SHB_C2[2] * 2.f * -y * SHV(6) + SHB_C2[4] * 2.f * -y * SHV(8);
              rz += SHB_C2[1] * y * SHV(5) + SHB_C2[2] * 4.f * z * SHV(6) + SHB_C2[3] * x * SHV(7);
              if (D > 2) {
                rx += SHB_C3[0] * SHV(9) * 6.f * xy + SHB_C3[1] * SHV(10) * yz + SHB_C3[2] * SHV(11) * -2.f * xy +
                      SHB_C3[3] * SHV(12) * -6.f * xz + SHB_C3[4] * SHV(13) * (-3.f * xx + 4.f * zz - yy) +
                      SHB_C3[5] * SHV(14) * 2.f * xz + SHB_C3[6] * SHV(15) * 3.f * (xx - yy);
                ry += SHB_C3[0] * SHV(9) * 3.f * (xx - yy) + SHB_C3[1] * SHV(10) * xz +
                      SHB_C3[2] * SHV(11) * (-3.f * yy + 4.f * zz - xx) + SHB_C3[3] * SHV(12) * -6.f * yz +
                      SHB_C3[4] * SHV(13) * -2.f * xy + SHB_C3[5] * SHV(14) * -2.f * yz + SHB_C3[6] * SHV(15) * -6.f * xy;
                rz += SHB_C3[1] * SHV(10) * xy + SHB_C3[2] * SHV(11) * 8.f * yz +
                      SHB_C3[3] * SHV(12) * 3.f * (2.f * zz - xx - yy) + SHB_C3[4] * SHV(13) * 8.f * xz +
                      SHB_C3[5] * SHV(14) * (xx - yy);
              }
            }
          }
#undef SHV
          ddx += rx * gch; ddy += ry * gch; ddz += rz * gch;
        }
        }
        sh_first = false;
        const float is32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
        dmx += ((sum2 - ox * ox) * ddx - oy * ox * ddy - oz * ox * ddz) * is32;
        dmy += (-ox * oy * ddx + (sum2 - oy * oy) * ddy - oz * oy * ddz) * is32;
        dmz += (-ox * oz * ddx - oy * oz * ddy + (sum2 - oz * oz) * ddz) * is32;
      } else {
        dcol[0] += dcr; dcol[1] += dcg; dcol[2] += dcb;
      }
    }
    if (go.dmeans2D) {
      float *o = go.dmeans2D + 3 * e;
      o[0] = g2x; o[1] = g2y; o[2] = 0.f;
    }
    if (go.dconic) {
      *reinterpret_cast<float4 *>(go.dconic + 4 * e) = make_float4(gA, gB, 0.f, gC);
    }
  }

  const bool accum = go.accumulate != 0;
  auto put = [&](float *p, float val) { if (accum) *p += val; else *p = val; };
  if (kVecSH && dsh_out) {
    // the whole row at once: the view sum, zeros above the active degree and for a Gaussian no view saw
    float4 *dsh4 = reinterpret_cast<float4 *>(dsh_out);
    const int nq = (M * 3) / 4;
#pragma unroll
    for (int j = 0; j < 12; j++) {
      if (j < nq) {
        float4 o = make_float4(s_dsh[kVecSH ? 4 * j : 0][threadIdx.x], s_dsh[kVecSH ? 4 * j + 1 : 0][threadIdx.x],
                               s_dsh[kVecSH ? 4 * j + 2 : 0][threadIdx.x], s_dsh[kVecSH ? 4 * j + 3 : 0][threadIdx.x]);
        if (accum) {
          const float4 c = dsh4[j];
          o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
        }
        dsh4[j] = o;
      }
    }
  } else if (HAS_SH && dsh_out && sh_first && !accum) {
    for (int k = 0; k < M * 3; k++) dsh_out[k] = 0.f;       // never visible in any view
  } else if (HAS_SH && dsh_out && !accum) {
    const int nb = (D + 1) * (D + 1);
    for (int k = nb * 3; k < M * 3; k++) dsh_out[k] = 0.f;  // coefficients above the active degree
  }
  if (go.dmeans3D) { put(go.dmeans3D + 3 * i, dmx); put(go.dmeans3D + 3 * i + 1, dmy); put(go.dmeans3D + 3 * i + 2, dmz); }
  if (go.dopacity) put(go.dopacity + i, dop);
  if (go.dcolors && !HAS_SH) { put(go.dcolors + 3 * i, dcol[0]); put(go.dcolors + 3 * i + 1, dcol[1]); put(go.dcolors + 3 * i + 2, dcol[2]); }
  if (go.dcov3D) {
#pragma unroll
    for (int k = 0; k < 6; k++) put(go.dcov3D + 6 * (size_t)i + k, dc3[k]);
  }
  if (from_sr && (go.dscales || go.drots)) {
    // ---- A.8 cov3D backward: Sigma = A A^T, A = R diag(s) ----
    const float G[3][3] = {{dc3[0], 0.5f * dc3[1], 0.5f * dc3[2]},
                           {0.5f * dc3[1], dc3[3], 0.5f * dc3[4]},
                           {0.5f * dc3[2], 0.5f * dc3[4], dc3[5]}};
    float dA[3][3];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int k = 0; k < 3; k++)
        dA[r][k] = 2.f * (G[r][0] * R[0][k] * s[k] + G[r][1] * R[1][k] * s[k] + G[r][2] * R[2][k] * s[k]);
    if (go.dscales) {
#pragma unroll
      for (int k = 0; k < 3; k++)
        put(go.dscales + 3 * i + k, R[0][k] * dA[0][k] + R[1][k] * dA[1][k] + R[2][k] * dA[2][k]);
    }
    if (go.drots) {
      float Dm[3][3];
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int k = 0; k < 3; k++) Dm[r][k] = dA[r][k] * s[k];
      const float4 q = reinterpret_cast<const float4 *>(g.rotations)[i];
      const float r = q.x, x = q.y, y = q.z, z = q.w;
      float4 dq;
      dq.x = 2.f * z * (Dm[1][0] - Dm[0][1]) + 2.f * y * (Dm[0][2] - Dm[2][0]) + 2.f * x * (Dm[2][1] - Dm[1][2]);
      dq.y = 2.f * y * (Dm[0][1] + Dm[1][0]) + 2.f * z * (Dm[0][2] + Dm[2][0]) + 2.f * r * (Dm[2][1] - Dm[1][2]) - 4.f * x * (Dm[1][1] + Dm[2][2]);
      dq.z = 2.f * x * (Dm[0][1] + Dm[1][0]) + 2.f * r * (Dm[0][2] - Dm[2][0]) + 2.f * z * (Dm[1][2] + Dm[2][1]) - 4.f * y * (Dm[0][0] + Dm[2][2]);
      dq.w = 2.f * r * (Dm[1][0] - Dm[0][1]) + 2.f * x * (Dm[0][2] + Dm[2][0]) + 2.f * y * (Dm[1][2] + Dm[2][1]) - 4.f * z * (Dm[0][0] + Dm[1][1]);
      float4 *o = reinterpret_cast<float4 *>(go.drots) + i;
      if (accum) { float4 c = *o; dq.x += c.x; dq.y += c.y; dq.z += c.z; dq.w += c.w; }
      *o = dq;
    }
  }
}

}  // namespace

cudaError_t launch_preprocess_backward(const GhrDims &d, const Layout &L, const Cameras &cam,
                                       const Gaussians &g, float scale_modifier, const char *state,
                                       const float *acc, const GradOut &go, cudaStream_t s) {
  if (d.P == 0) return cudaSuccess;
  constexpr int kT = GHR_PBWD_THREADS;          // (build-time A/B: 64 balances 469 blocks over 148 SMs better)
  int nb = (d.P + kT - 1) / kT;
  const float4 *geom = (const float4 *)(state + L.pub.off_geom);
  const uint8_t *cl = (const uint8_t *)(state + L.pub.off_clamped);
  const bool has_sh = g.shs && !g.colors_precomp;
  const bool vec_sh = has_sh && (d.M * 3) % 4 == 0 && ((uintptr_t)g.shs & 15) == 0 &&
                      (go.dsh == nullptr || ((uintptr_t)go.dsh & 15) == 0);
  if (vec_sh)
    preprocess_backward_kernel<true, true><<<nb, kT, 0, s>>>(d.P, d.V, d.H, d.W, d.M, d.sh_degree, scale_modifier,
                                                              cam, g, geom, cl, acc, go);
  else if (has_sh)
    preprocess_backward_kernel<true, false><<<nb, kT, 0, s>>>(d.P, d.V, d.H, d.W, d.M, d.sh_degree, scale_modifier,
                                                               cam, g, geom, cl, acc, go);
  else
    preprocess_backward_kernel<false, false><<<nb, kT, 0, s>>>(d.P, d.V, d.H, d.W, d.M, d.sh_degree, scale_modifier,
                                                                cam, g, geom, cl, acc, go);
  return cudaGetLastError();
}

}  // namespace ghr
