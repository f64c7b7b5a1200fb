// Fused Gaussian-attribute head: activations + interaction-aware attribute blending, forward and
// backward, one thread per Gaussian.
//
// Replaces, in one launch each way, the ~15 elementwise PyTorch kernels the reference runs per view
// between its Linear heads and the rasterizer call:
//   GSLayer.forward            /root/reference/tgs/models/renderer_one_shot.py:191-214
//       rotation -> F.normalize, scaling -> trunc_exp (+ clamp), opacity -> sigmoid,
//       shs (use_rgb) -> sigmoid, xyz -> (restricted) offset + pts
//   forward_single_view        renderer_one_shot.py:298-334
//       means3D += xyz_b, opacity += opacity_b, colour = colour*w0 + w1 - 1 (+ b0)
//   trunc_exp                  /root/reference/tgs/utils/ops.py:37-53  (backward: g * exp(min(x, 15)))
// Outputs land in the layout ghr_forward takes (means3D, scales, rotations, opacities, colours).
#include "ghr_internal.cuh"

namespace ghr {

namespace {

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

constexpr float kMaxStep = 1.2f / 32.0f;      // renderer_one_shot.py:208
constexpr float kNormEps = 1e-12f;            // torch.nn.functional.normalize default eps

__global__ void __launch_bounds__(256)
attributes_forward_kernel(GhrAttributeArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.P) return;
  const bool offset = a.flags & GHR_ATTR_XYZ_OFFSET, restrict_ = a.flags & GHR_ATTR_RESTRICT_OFFSET;
  const bool clip = a.flags & GHR_ATTR_CLIP_SCALING;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    // xyz (:205-210) + xyz_b (:300-301)
    float v = a.xyz_raw[3 * i + c];
    if (restrict_) v = (sigmoidf(v) - 0.5f) * kMaxStep;
    float m = offset ? v + a.pts[3 * i + c] : a.pts[3 * i + c];
    if (a.xyz_b) m += a.xyz_b[3 * i + c];
    a.means3D[3 * i + c] = m;
    // scaling (:196-199)
    float s = expf(a.scaling_raw[3 * i + c]);
    if (clip) s = fminf(fmaxf(s, 0.0f), a.clip_scaling);
    a.scales[3 * i + c] = s;
    // colour, use_rgb path (:201-204, :321-328); the SH path (rgb_raw == NULL) blends its coefficients with
    // ghr_sh_blend_forward
    if (a.rgb_raw) {
      float col = sigmoidf(a.rgb_raw[3 * i + c]);
      if (a.color_w0) col = col * a.color_w0[3 * i + c] + a.color_w1[3 * i + c] - 1.0f;
      if (a.color_b0) col += a.color_b0[3 * i + c];
      a.colors[3 * i + c] = col;
    }
  }
  // rotation (:194-195)
  const float4 r = reinterpret_cast<const float4 *>(a.rotation_raw)[i];
  const float n = fmaxf(sqrtf(r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w), kNormEps);
  reinterpret_cast<float4 *>(a.rotations)[i] = make_float4(r.x / n, r.y / n, r.z / n, r.w / n);
  // opacity (:200-201) + opacity_b (:306-307)
  float o = sigmoidf(a.opacity_raw[i]);
  if (a.opacity_b) o += a.opacity_b[i];
  a.opacities[i] = o;
}

__global__ void __launch_bounds__(256)
attributes_backward_kernel(GhrAttributeArgs a, GhrAttributeGrads g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.P) return;
  const bool offset = a.flags & GHR_ATTR_XYZ_OFFSET, restrict_ = a.flags & GHR_ATTR_RESTRICT_OFFSET;
  const bool clip = a.flags & GHR_ATTR_CLIP_SCALING;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const int k = 3 * i + c;
    const float gm = g.dL_dmeans3D ? g.dL_dmeans3D[k] : 0.0f;
    if (g.d_pts) g.d_pts[k] = gm;
    if (g.d_xyz_b) g.d_xyz_b[k] = gm;
    if (g.d_xyz_raw) {
      float d = 0.0f;
      if (offset) {
        d = gm;
        if (restrict_) {
          const float sg = sigmoidf(a.xyz_raw[k]);
          d = gm * sg * (1.0f - sg) * kMaxStep;
        }
      }
      g.d_xyz_raw[k] = d;
    }
    if (g.d_scaling_raw) {
      const float x = a.scaling_raw[k];
      float gs = g.dL_dscales ? g.dL_dscales[k] : 0.0f;
      if (clip) {
        const float s = expf(x);
        if (!(s >= 0.0f && s <= a.clip_scaling)) gs = 0.0f;   // torch.clamp passes the gradient inside [min, max]
      }
      g.d_scaling_raw[k] = gs * expf(fminf(x, 15.0f));          // trunc_exp backward (ops.py:50-53)
    }
    if (a.rgb_raw) {
      const float gc = g.dL_dcolors ? g.dL_dcolors[k] : 0.0f;
      const float sg = sigmoidf(a.rgb_raw[k]);
      const float w0 = a.color_w0 ? a.color_w0[k] : 1.0f;
      if (g.d_rgb_raw) g.d_rgb_raw[k] = gc * w0 * sg * (1.0f - sg);
      if (g.d_color_w0) g.d_color_w0[k] = gc * sg;
      if (g.d_color_w1) g.d_color_w1[k] = gc;
      if (g.d_color_b0) g.d_color_b0[k] = gc;
    }
  }
  if (g.d_rotation_raw) {
    const float4 r = reinterpret_cast<const float4 *>(a.rotation_raw)[i];
    float4 gq = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g.dL_drotations) gq = reinterpret_cast<const float4 *>(g.dL_drotations)[i];
    const float len = sqrtf(r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w);
    float4 d;
    if (len > kNormEps) {
      // q = r/|r|:  dL/dr = (g - q (q.g)) / |r|
      const float inv = 1.0f / len;
      const float4 q = make_float4(r.x * inv, r.y * inv, r.z * inv, r.w * inv);
      const float qg = q.x * gq.x + q.y * gq.y + q.z * gq.z + q.w * gq.w;
      d = make_float4((gq.x - q.x * qg) * inv, (gq.y - q.y * qg) * inv, (gq.z - q.z * qg) * inv,
                      (gq.w - q.w * qg) * inv);
    } else {
      d = make_float4(gq.x / kNormEps, gq.y / kNormEps, gq.z / kNormEps, gq.w / kNormEps);   // clamp_min branch
    }
    reinterpret_cast<float4 *>(g.d_rotation_raw)[i] = d;
  }
  const float go = g.dL_dopacity ? g.dL_dopacity[i] : 0.0f;
  if (g.d_opacity_raw) {
    const float sg = sigmoidf(a.opacity_raw[i]);
    g.d_opacity_raw[i] = go * sg * (1.0f - sg);
  }
  if (g.d_opacity_b) g.d_opacity_b[i] = go;
}

// SH path of the blending (renderer_one_shot.py:329-334): 4 coefficients per thread when everything is 16-byte
// aligned, else one
template <int kVec>
__global__ void __launch_bounds__(256)
sh_blend_forward_kernel(int64_t n, const float *__restrict__ x, const float *__restrict__ w,
                        const float *__restrict__ b, float *__restrict__ out) {
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * kVec;
  if (i0 >= n) return;
  float xv[kVec], wv[kVec], bv[kVec], o[kVec];
  if (kVec == 4) {
    *reinterpret_cast<float4 *>(xv) = *reinterpret_cast<const float4 *>(x + i0);
    if (w) *reinterpret_cast<float4 *>(wv) = *reinterpret_cast<const float4 *>(w + i0);
    if (b) *reinterpret_cast<float4 *>(bv) = *reinterpret_cast<const float4 *>(b + i0);
  } else {
    xv[0] = x[i0];
    if (w) wv[0] = w[i0];
    if (b) bv[0] = b[i0];
  }
#pragma unroll
  for (int k = 0; k < kVec; k++) {
    float v = xv[k];
    if (w) v = __fmul_rn(v, wv[k]);                    // shs * color_w                       (:331-332)
    if (b) v = __fadd_rn(__fmul_rn(v, wv[k]), bv[k]);  // (shs * color_w) * color_w + color_b (:333-334)
    o[k] = v;
  }
  if (kVec == 4) *reinterpret_cast<float4 *>(out + i0) = *reinterpret_cast<float4 *>(o);
  else out[i0] = o[0];
}

template <int kVec>
__global__ void __launch_bounds__(256)
sh_blend_backward_kernel(int64_t n, const float *__restrict__ x, const float *__restrict__ w,
                         const float *__restrict__ b, const float *__restrict__ go, float *__restrict__ d_x,
                         float *__restrict__ d_w, float *__restrict__ d_b) {
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * kVec;
  if (i0 >= n) return;
#pragma unroll
  for (int k = 0; k < kVec; k++) {
    const int64_t i = i0 + k;
    const float g = go[i], xv = x[i], wv = w ? w[i] : 1.0f;
    // out = x w (w only) or x w^2 + b (w and b)
    if (d_x) d_x[i] = b ? g * wv * wv : g * wv;
    if (d_w) d_w[i] = b ? 2.0f * g * xv * wv : g * xv;
    if (d_b) d_b[i] = g;
  }
}

int check_attr(const char *fn, const GhrAttributeArgs *a) {
  if (!a) { set_error("%s: NULL args", fn); return GHR_EINVAL; }
  if (a->P < 0) { set_error("%s: negative P", fn); return GHR_EINVAL; }
  if (a->P > 0 && (!a->xyz_raw || !a->pts || !a->scaling_raw || !a->rotation_raw || !a->opacity_raw)) {
    set_error("%s: xyz_raw/pts/scaling_raw/rotation_raw/opacity_raw are required", fn);
    return GHR_EINVAL;
  }
  if (!a->rgb_raw && (a->color_w0 || a->color_b0)) {
    set_error("%s: colour blending terms without rgb_raw (the SH path blends with ghr_sh_blend_forward)", fn);
    return GHR_EINVAL;
  }
  if ((a->color_w0 != nullptr) != (a->color_w1 != nullptr)) {
    set_error("%s: color_w0 and color_w1 come together", fn);
    return GHR_EINVAL;
  }
  return GHR_OK;
}

}  // namespace

}  // namespace ghr

using namespace ghr;

extern "C" {

int ghr_attributes_forward(const GhrAttributeArgs *a, void *cuda_stream) {
  int rc = check_attr("ghr_attributes_forward", a);
  if (rc != GHR_OK) return rc;
  if (a->P == 0) return GHR_OK;
  if (!a->means3D || !a->scales || !a->rotations || !a->opacities || (a->rgb_raw && !a->colors)) {
    set_error("ghr_attributes_forward: all five outputs are required (colors only with rgb_raw)");
    return GHR_EINVAL;
  }
  attributes_forward_kernel<<<(a->P + 255) / 256, 256, 0, (cudaStream_t)cuda_stream>>>(*a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("ghr_attributes_forward: %s", cudaGetErrorString(e)); return GHR_ECUDA; }
  return GHR_OK;
}

int ghr_attributes_backward(const GhrAttributeArgs *a, const GhrAttributeGrads *g, void *cuda_stream) {
  int rc = check_attr("ghr_attributes_backward", a);
  if (rc != GHR_OK) return rc;
  if (!g) { set_error("ghr_attributes_backward: NULL grads"); return GHR_EINVAL; }
  if (a->P == 0) return GHR_OK;
  attributes_backward_kernel<<<(a->P + 255) / 256, 256, 0, (cudaStream_t)cuda_stream>>>(*a, *g);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("ghr_attributes_backward: %s", cudaGetErrorString(e)); return GHR_ECUDA; }
  return GHR_OK;
}

int ghr_sh_blend_forward(int64_t n, const float *x, const float *w, const float *b, float *out, void *cuda_stream) {
  if (n < 0 || (n > 0 && (!x || !out))) { set_error("ghr_sh_blend_forward: bad arguments"); return GHR_EINVAL; }
  if (b && !w) { set_error("ghr_sh_blend_forward: color_b needs color_w (renderer_one_shot.py:333-334)"); return GHR_EINVAL; }
  if (n == 0) return GHR_OK;
  const bool vec = n % 4 == 0 && (((uintptr_t)x | (uintptr_t)out | (uintptr_t)w | (uintptr_t)b) & 15) == 0;
  cudaStream_t s = (cudaStream_t)cuda_stream;
  if (vec) sh_blend_forward_kernel<4><<<(unsigned)((n / 4 + 255) / 256), 256, 0, s>>>(n, x, w, b, out);
  else sh_blend_forward_kernel<1><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, x, w, b, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("ghr_sh_blend_forward: %s", cudaGetErrorString(e)); return GHR_ECUDA; }
  return GHR_OK;
}

int ghr_sh_blend_backward(int64_t n, const float *x, const float *w, const float *b, const float *dL_dout, float *d_x,
                          float *d_w, float *d_b, void *cuda_stream) {
  if (n < 0 || (n > 0 && (!x || !dL_dout))) { set_error("ghr_sh_blend_backward: bad arguments"); return GHR_EINVAL; }
  if ((b && !w) || (d_w && !w) || (d_b && !b)) { set_error("ghr_sh_blend_backward: gradient of an absent term"); return GHR_EINVAL; }
  if (n == 0) return GHR_OK;
  sh_blend_backward_kernel<1><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(n, x, w, b, dL_dout, d_x,
                                                                                                  d_w, d_b);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("ghr_sh_blend_backward: %s", cudaGetErrorString(e)); return GHR_ECUDA; }
  return GHR_OK;
}

}  // extern "C"
