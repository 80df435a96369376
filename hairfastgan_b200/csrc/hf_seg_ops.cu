// Kernels around the tcgen05 convolution for the BiSeNet face-parsing network (reference
// models/CtrlHair/external_code/face_parsing/{model,resnet}.py): the 7x7 stride-2 RGB stem, 3x3 stride-2 max pooling,
// pooled 1x1 "attention" convolutions, gated add with nearest 2x upsampling, and the final bilinear logit upsampling.
// All SIMT, fp32 math, deterministic.
#include <algorithm>

#include "hf_kernels.cuh"

namespace hf {

static inline int cdiv_s(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

template <int DT>
__device__ __forceinline__ void unpack8s(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    f[2 * k] = Half2T<DT>::to_float((uint16_t)(w[k] & 0xFFFF));
    f[2 * k + 1] = Half2T<DT>::to_float((uint16_t)(w[k] >> 16));
  }
}
template <int DT>
__device__ __forceinline__ uint4 pack8s(const float* v) {
  uint4 o;
  o.x = Half2T<DT>::pack(v[0], v[1]); o.y = Half2T<DT>::pack(v[2], v[3]);
  o.z = Half2T<DT>::pack(v[4], v[5]); o.w = Half2T<DT>::pack(v[6], v[7]);
  return o;
}

// ------------------------------------------------------------------------------------------------
// Resnet18.conv1 + bn1 + relu (resnet.py:60-61,69-70): 7x7, stride 2, pad 3, 3 -> 64 channels, fp32 NCHW in,
// 16-bit NHWC out.  w = [147][64] fp32 (tap-major, BatchNorm scale folded), shift[64].
// CTA = 16x16 output pixels: the 37x37x3 input patch and the weights live in shared memory; one thread = one pixel
// x 64 channels (64 accumulators; weights are broadcast LDS.128).  K = 147 is too thin for the tensor-core path
// (Cin = 3 would be padded to 32 per tap: 11x wasted MMA work) and the layer is 1.2 GFLOP per 512^2 image.
// ------------------------------------------------------------------------------------------------
constexpr int kStemT = 16, kStemP = 2 * kStemT + 5;      // 37

template <int DT>
__global__ void __launch_bounds__(256) stem7x7_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ shift, uint16_t* __restrict__ y, int H,
                                                      int W, int Ho, int Wo) {
  extern __shared__ float sm[];
  float* sw = sm;                              // [147][64]
  float* sx = sm + 147 * 64;                   // [3][37][37]
  const int b = blockIdx.z, oy0 = blockIdx.y * kStemT, ox0 = blockIdx.x * kStemT;
  for (int i = threadIdx.x; i < 147 * 64; i += 256) sw[i] = __ldg(w + i);
  const int iy0 = 2 * oy0 - 3, ix0 = 2 * ox0 - 3;
  for (int i = threadIdx.x; i < 3 * kStemP * kStemP; i += 256) {
    const int c = i / (kStemP * kStemP), r = (i / kStemP) % kStemP, q = i % kStemP;
    const int yy = iy0 + r, xx = ix0 + q;
    sx[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(x + (((size_t)b * 3 + c) * H + yy) * W + xx) : 0.f;
  }
  __syncthreads();
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[64];
#pragma unroll
  for (int o = 0; o < 64; ++o) acc[o] = 0.f;
#pragma unroll 1
  for (int c = 0; c < 3; ++c) {
#pragma unroll 1
    for (int ky = 0; ky < 7; ++ky) {
      const float* row = sx + (c * kStemP + 2 * ty + ky) * kStemP + 2 * tx;
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const float v = row[kx];
        const float4* wp = reinterpret_cast<const float4*>(sw + ((c * 7 + ky) * 7 + kx) * 64);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float4 ww = wp[q];
          acc[4 * q] = fmaf(v, ww.x, acc[4 * q]); acc[4 * q + 1] = fmaf(v, ww.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(v, ww.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(v, ww.w, acc[4 * q + 3]);
        }
      }
    }
  }
  const int oy = oy0 + ty, ox = ox0 + tx;
  if (oy < Ho && ox < Wo) {
    uint4* dst = reinterpret_cast<uint4*>(y + (((size_t)b * Ho + oy) * Wo + ox) * 64);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = fmaxf(acc[g * 8 + k] + __ldg(shift + g * 8 + k), 0.f);
      dst[g] = pack8s<DT>(v);
    }
  }
}

int launch_stem7x7(const float* x, const float* w, const float* shift, void* y16, int B, int H, int W, int dtype,
                   cudaStream_t st) {
  HF_REQUIRE(x && w && shift && y16, "stem7x7: null pointer");
  HF_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, "stem7x7: bad shape");
  const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
  const size_t smem = (size_t)(147 * 64 + 3 * kStemP * kStemP) * sizeof(float);
  dim3 grid(cdiv_s(Wo, kStemT), cdiv_s(Ho, kStemT), B);
  if (dtype == HF_BF16) {
    HF_CUDA_OK(cudaFuncSetAttribute(stem7x7_kernel<HF_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stem7x7_kernel<HF_BF16><<<grid, 256, smem, st>>>(x, w, shift, (uint16_t*)y16, H, W, Ho, Wo);
  } else {
    HF_CUDA_OK(cudaFuncSetAttribute(stem7x7_kernel<HF_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stem7x7_kernel<HF_F16><<<grid, 256, smem, st>>>(x, w, shift, (uint16_t*)y16, H, W, Ho, Wo);
  }
  HF_LAUNCH_OK("stem7x7");
  count_launch();
  return HF_OK;
}

// ------------------------------------------------------------------------------------------------
// nn.MaxPool2d(kernel_size=3, stride=2, padding=1) (resnet.py:62,71) on NHWC 16-bit; padding never wins the max.
// ------------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y,
                                                           int H, int W, int Ho, int Wo, int C8, int64_t total8) {
  const int C = C8 * 8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) * 8;
    int64_t t = i / C8;
    const int ox = (int)(t % Wo); t /= Wo;
    const int oy = (int)(t % Ho);
    const int b = (int)(t / Ho);
    float m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = -3.0e38f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = 2 * oy + dy;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = 2 * ox + dx;
        if (xx < 0 || xx >= W) continue;
        float v[8];
        unpack8s<DT>(__ldg(reinterpret_cast<const uint4*>(x + (((size_t)b * H + yy) * W + xx) * C + c)), v);
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], v[k]);
      }
    }
    *reinterpret_cast<uint4*>(y + i * 8) = pack8s<DT>(m);
  }
}

int launch_maxpool3x3s2(const void* x16, void* y16, int B, int H, int W, int C, int dtype, cudaStream_t st) {
  HF_REQUIRE(x16 && y16 && C % 8 == 0 && B > 0 && H > 0 && W > 0, "maxpool: bad arguments");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const int64_t total8 = (int64_t)B * Ho * Wo * C / 8;
  const int grid = (int)std::min<int64_t>((total8 + 255) / 256, (int64_t)num_sms() * 16);
  if (dtype == HF_BF16)
    maxpool3x3s2_kernel<HF_BF16><<<grid, 256, 0, st>>>((const uint16_t*)x16, (uint16_t*)y16, H, W, Ho, Wo, C / 8, total8);
  else
    maxpool3x3s2_kernel<HF_F16><<<grid, 256, 0, st>>>((const uint16_t*)x16, (uint16_t*)y16, H, W, Ho, Wo, C / 8, total8);
  HF_LAUNCH_OK("maxpool3x3s2");
  count_launch();
  return HF_OK;
}

// ------------------------------------------------------------------------------------------------
// 1x1 convolution of the globally pooled feature (+ folded BatchNorm + activation):
//   out[b,o] = act((sum_c w[o,c] * mean_hw(x)[b,c]) * scale[o] + shift[o])
// AttentionRefinementModule attention (model.py:82-86: sigmoid) and ContextPath.conv_avg (model.py:114-115: ReLU).
// Stage 1 is the shared deterministic pooling (channel_sum_partial_kernel); this is stage 2, one CTA per sample.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pooled_fc_kernel(const float* __restrict__ part, int S, float inv_hw,
                                                        const float* __restrict__ w, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, int act,
                                                        float* __restrict__ out, int C, int Cout) {
  extern __shared__ float mean[];
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int c = tid; c < C; c += 256) {
    float t = 0.f;
    for (int s = 0; s < S; ++s) t += part[((size_t)b * S + s) * C + c];
    mean[c] = t * inv_hw;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int o = warp; o < Cout; o += 8) {
    float t = 0.f;
    for (int c = lane; c < C; c += 32) t = fmaf(__ldg(w + (size_t)o * C + c), mean[c], t);
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, k);
    if (lane == 0) {
      float v = fmaf(t, scale ? __ldg(scale + o) : 1.f, shift ? __ldg(shift + o) : 0.f);
      if (act == 2) v = 1.f / (1.f + __expf(-v));
      else if (act == 1) v = fmaxf(v, 0.f);
      out[(size_t)b * Cout + o] = v;
    }
  }
}

int launch_pooled_fc(const void* x16, const float* w, const float* scale, const float* shift, int act, float* out,
                     float* ws, int B, int HW, int C, int Cout, int dtype, cudaStream_t st) {
  HF_REQUIRE(x16 && w && out && ws, "pooled_fc: null pointer");
  HF_REQUIRE(C % 8 == 0 && C <= 8192 && Cout > 0 && act >= 0 && act <= 2, "pooled_fc: bad arguments");
  int S = 1;
  int rc = launch_channel_partial(x16, ws, B, HW, C, dtype, st, &S);
  if (rc) return rc;
  pooled_fc_kernel<<<B, 256, (size_t)C * sizeof(float), st>>>(ws, S, 1.f / (float)HW, w, scale, shift, act, out, C, Cout);
  HF_LAUNCH_OK("pooled_fc");
  count_launch();
  return HF_OK;
}

// ------------------------------------------------------------------------------------------------
// y[b,Y,X,c] = x[b,Y/up,X/up,c] * gate[b,c] + addvec[b,c] + addt[b,Y/up,X/up,c]      (up = 1 or 2, nearest)
// ContextPath (model.py:116-128): feat*atten (+ the pooled branch | + the upsampled coarser branch), then
// F.interpolate(mode='nearest') to the next finer level.
// ------------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(256) gate_add_up_kernel(const uint16_t* __restrict__ x, const float* __restrict__ gate,
                                                          const float* __restrict__ addvec,
                                                          const uint16_t* __restrict__ addt, uint16_t* __restrict__ y,
                                                          int h, int w, int up, int C8, int64_t total8) {
  const int C = C8 * 8, H = h * up, W = w * up;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) * 8;
    int64_t t = i / C8;
    const int X = (int)(t % W); t /= W;
    const int Y = (int)(t % H);
    const int b = (int)(t / H);
    const size_t src = (((size_t)b * h + Y / up) * w + X / up) * C + c;
    float v[8];
    unpack8s<DT>(__ldg(reinterpret_cast<const uint4*>(x + src)), v);
    if (gate) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] *= __ldg(gate + (size_t)b * C + c + k);
    }
    if (addvec) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += __ldg(addvec + (size_t)b * C + c + k);
    }
    if (addt) {
      float a[8];
      unpack8s<DT>(__ldg(reinterpret_cast<const uint4*>(addt + src)), a);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += a[k];
    }
    *reinterpret_cast<uint4*>(y + i * 8) = pack8s<DT>(v);
  }
}

int launch_gate_add_up(const void* x16, const float* gate, const float* addvec, const void* addt16, void* y16, int B,
                       int h, int w, int C, int up, int dtype, cudaStream_t st) {
  HF_REQUIRE(x16 && y16 && C % 8 == 0 && (up == 1 || up == 2) && B > 0 && h > 0 && w > 0, "gate_add_up: bad arguments");
  const int64_t total8 = (int64_t)B * h * up * w * up * C / 8;
  const int grid = (int)std::min<int64_t>((total8 + 255) / 256, (int64_t)num_sms() * 16);
  if (dtype == HF_BF16)
    gate_add_up_kernel<HF_BF16><<<grid, 256, 0, st>>>((const uint16_t*)x16, gate, addvec, (const uint16_t*)addt16,
                                                      (uint16_t*)y16, h, w, up, C / 8, total8);
  else
    gate_add_up_kernel<HF_F16><<<grid, 256, 0, st>>>((const uint16_t*)x16, gate, addvec, (const uint16_t*)addt16,
                                                     (uint16_t*)y16, h, w, up, C / 8, total8);
  HF_LAUNCH_OK("gate_add_up");
  count_launch();
  return HF_OK;
}

// ------------------------------------------------------------------------------------------------
// F.interpolate(x, (H, W), mode='bilinear', align_corners=True) on fp32 NCHW (model.py:239-241): the first C of
// Cin channel planes of x (the logit convolution pads its 19 classes to 32 output channels).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bilinear_up_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, int C,
                                                               int Cin, int h, int w, int H, int W, int64_t total) {
  const float ry = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, rx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int X = (int)(i % W);
    int64_t t = i / W;
    const int Y = (int)(t % H); t /= H;
    const int c = (int)(t % C);
    const int b = (int)(t / C);
    const float fy = Y * ry, fx = X * rx;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + 1 < h ? y0 + 1 : y0, x1 = x0 + 1 < w ? x0 + 1 : x0;
    const float ly = fy - y0, lx = fx - x0;
    const float* p = x + ((size_t)b * Cin + c) * h * w;
    const float a00 = __ldg(p + (size_t)y0 * w + x0), a01 = __ldg(p + (size_t)y0 * w + x1);
    const float a10 = __ldg(p + (size_t)y1 * w + x0), a11 = __ldg(p + (size_t)y1 * w + x1);
    y[i] = (1.f - ly) * ((1.f - lx) * a00 + lx * a01) + ly * ((1.f - lx) * a10 + lx * a11);
  }
}

int launch_bilinear_up_nchw(const float* x, float* y, int B, int C, int Cin, int h, int w, int H, int W,
                            cudaStream_t st) {
  HF_REQUIRE(x && y && B > 0 && C > 0 && Cin >= C && h > 0 && w > 0 && H > 0 && W > 0, "bilinear_up: bad arguments");
  const int64_t total = (int64_t)B * C * H * W;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 32);
  bilinear_up_nchw_kernel<<<grid, 256, 0, st>>>(x, y, C, Cin, h, w, H, W, total);
  HF_LAUNCH_OK("bilinear_up_nchw");
  count_launch();
  return HF_OK;
}

}  // namespace hf
