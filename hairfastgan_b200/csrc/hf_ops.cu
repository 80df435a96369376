// HBM-bound SIMT kernels of the hot path: upfirdn2d, fused bias+activation, style affine /
// demodulation tables, NCHW->NHWC modulate+cast pre-pass, RGB combine (+ up-FIR of the skip),
// standalone ToRGB, weight packing.  All fp32 math; 16-bit only as storage for tensor-core operands.
#include "hf_kernels.cuh"
#include <type_traits>

namespace hf {

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// =============================================================================================
// upfirdn2d  (reference semantics: models/stylegan2/op/upfirdn2d.py:159-200,
//             op/upfirdn2d_kernel.cu:107-207 -- zero-stuff, pad/crop, convolve (flipped taps), decimate)
// =============================================================================================

// Fast path: up = 1, 4x4 taps, down in {1,2}.  One CTA = one (plane, 32x64 | 16x64 output tile);
// the input window is staged once in shared memory with coalesced row loads, each thread produces
// 4 consecutive outputs per row from a sliding register window and stores them as one float4.
template <int DOWN, int ROWS>
__global__ void __launch_bounds__(256) upfirdn2d_up1_k4_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                               const float* __restrict__ k, int in_h, int in_w,
                                                               int out_h, int out_w, int px0, int py0) {
  constexpr int TW = 64;
  constexpr int IN_ROWS = (ROWS - 1) * DOWN + 4;
  constexpr int IN_COLS = (TW - 1) * DOWN + 4;
  constexpr int PITCH = (IN_COLS + 3) / 4 * 4;          // rows start 16-byte aligned: the window is read as float4
  __shared__ __align__(16) float tile[IN_ROWS * PITCH];
  __shared__ float kf[16];
  const int plane = blockIdx.z;
  const int oy0 = blockIdx.y * ROWS, ox0 = blockIdx.x * TW;
  const float* xp = x + (size_t)plane * in_h * in_w;
  if (threadIdx.x < 16) kf[threadIdx.x] = k[15 - threadIdx.x];   // flipped: kf[ky][kx] = k[3-ky][3-kx]
  const int gy0 = oy0 * DOWN - py0, gx0 = ox0 * DOWN - px0;
  // staging in two unrolled phases -- all global loads of a thread are issued before the first shared store, so
  // their latencies overlap (a rolled load->store loop paid one HBM round trip per element: 0.29 of the copy peak)
  constexpr int N_STAGE = (IN_ROWS * IN_COLS + 255) / 256;
  float stage[N_STAGE];
#pragma unroll
  for (int it = 0; it < N_STAGE; ++it) {
    const int i = threadIdx.x + it * 256;
    const int r = i / IN_COLS, c = i - r * IN_COLS;
    const int gy = gy0 + r, gx = gx0 + c;
    stage[it] = (i < IN_ROWS * IN_COLS && gy >= 0 && gy < in_h && gx >= 0 && gx < in_w)
                    ? __ldg(xp + (size_t)gy * in_w + gx) : 0.f;
  }
#pragma unroll
  for (int it = 0; it < N_STAGE; ++it) {
    const int i = threadIdx.x + it * 256;
    if (i < IN_ROWS * IN_COLS) {
      const int r = i / IN_COLS, c = i - r * IN_COLS;
      tile[r * PITCH + c] = stage[it];
    }
  }
  __syncthreads();
  float kr[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) kr[i] = kf[i];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16 threads
  float* yp = y + (size_t)plane * out_h * out_w;
#pragma unroll
  for (int rr = 0; rr < ROWS / 16; ++rr) {
    const int r = ty + rr * 16;
    const int oy = oy0 + r, ox = ox0 + tx * 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const float4* row4 = reinterpret_cast<const float4*>(tile + (r * DOWN + ky) * PITCH + tx * 4 * DOWN);
      constexpr int NW4 = (3 * DOWN + 4 + 3) / 4;
      float w[NW4 * 4];
#pragma unroll
      for (int i = 0; i < NW4; ++i) {
        const float4 t4 = row4[i];
        w[4 * i] = t4.x; w[4 * i + 1] = t4.y; w[4 * i + 2] = t4.z; w[4 * i + 3] = t4.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) acc[j] = fmaf(w[j * DOWN + kx], kr[ky * 4 + kx], acc[j]);
    }
    if (oy < out_h) {
      float* dst = yp + (size_t)oy * out_w + ox;
      if (ox + 3 < out_w && ((out_w & 3) == 0)) {
        *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (ox + j < out_w) dst[j] = acc[j];
      }
    }
  }
}

// Fast path: up = 2, down = 1, 4x4 taps (the RGB-skip Upsample, model.py:35-53).  Only 2x2 of the
// 16 taps hit a non-zero sample; each thread makes 4 consecutive outputs of one row.  (A row-per-CTA variant with
// the two live input rows staged in shared memory measured slower: 1.06 vs 1.41 TB/s.)
__global__ void __launch_bounds__(256) upfirdn2d_up2_k4_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                               const float* __restrict__ k, int in_h, int in_w,
                                                               int out_h, int out_w, int px0, int py0,
                                                               int64_t total_quads) {
  __shared__ float kf[16];
  if (threadIdx.x < 16) kf[threadIdx.x] = k[15 - threadIdx.x];
  __syncthreads();
  const int quads_per_row = (out_w + 3) >> 2;
  // 32-bit index math (the launcher checks total_quads < 2^31)
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < (int)total_quads; q += gridDim.x * blockDim.x) {
    int qx = q % quads_per_row;
    int t = q / quads_per_row;
    int oy = t % out_h;
    int plane = t / out_h;
    const float* xp = x + (size_t)plane * in_h * in_w;
    const int ky0 = (py0 - oy) & 1;          // oy + ky - py0 must be even
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int ky = ky0 + 2 * a;
      const int uy = oy + ky - py0;
      const int iy = uy >> 1;
      if (uy < 0 || iy >= in_h) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ox = qx * 4 + j;
        const int kx0 = (px0 - ox) & 1;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int kx = kx0 + 2 * b;
          const int ux = ox + kx - px0;
          const int ix = ux >> 1;
          if (ux >= 0 && ix < in_w) acc[j] = fmaf(__ldg(xp + (size_t)iy * in_w + ix), kf[ky * 4 + kx], acc[j]);
        }
      }
    }
    float* dst = y + ((size_t)plane * out_h + oy) * out_w + qx * 4;
    if ((out_w & 3) == 0) {
      *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else {
      for (int j = 0; j < 4; ++j)
        if (qx * 4 + j < out_w) dst[j] = acc[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Round-2 fast paths: register sliding windows, no shared memory, no block-level synchronisation.
// A warp owns a strip of RS rows x 128 columns and walks down it; every lane keeps its 4 (or 2x8) outputs' partial
// sums in registers and each input row is fetched ONCE per strip (plus the 3 / 2 halo rows, which the neighbouring
// strip of the same CTA has just pulled into L1/L2).  All global accesses of a warp are row-contiguous (512 B per
// row); the 4x4 FIR costs 64 FMAs per float4 of output, well under the issue budget of an HBM-bound kernel.
// ---------------------------------------------------------------------------------------------
// up = 1, down = 1, 4x4 taps, any pad >= 0 (the Blur after the transposed convolution, model.py:77-93).
// Input side (round 2, after ncu: the register-window loads left every warp with ~1 row in flight -- long-scoreboard
// stalls 15 per issue, 0.65 of the copy peak): each warp owns a ring of kUpStages row segments in shared memory and
// keeps kUpStages - 1 rows AHEAD in flight with 4-byte cp.async (the (2R+1)-wide rows are only 4-byte aligned, so
// neither 16-byte cp.async nor TMA applies): lane l copies elements l, l+32, ... of the 131-float segment, fully
// coalesced, zero-filled outside the image (= the padding).  No block-level synchronisation: a warp only waits for
// its own copy groups.
constexpr int kUpStages = 8;                                 // power of two: slot = row & 7
constexpr int kUpSeg = 136;                                  // 128 + 3 taps, padded to a multiple of 8 floats

__device__ __forceinline__ void cp_async4_zfill(uint32_t dst_s, const float* src, bool pred) {
  const int sz = pred ? 4 : 0;                               // src-size 0: the 4 destination bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst_s), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int RS>
__global__ void __launch_bounds__(256) upfirdn2d_up1_k4_strip_kernel(const float* __restrict__ x,
                                                                     float* __restrict__ y,
                                                                     const float* __restrict__ k, int in_h, int in_w,
                                                                     int out_h, int out_w, int px0, int py0, int wx) {
  // wx (1, 2, 4 or 8) warps of a CTA sit SIDE BY SIDE (wx x 128 columns, up to a whole 1024-wide row) and walk down the
  // same RS rows; the other 8 / wx warp groups take the strips below
  __shared__ __align__(16) float ring[8][kUpStages][kUpSeg];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int oxw = (blockIdx.x * wx + (warp & (wx - 1))) * 128;       // first output column of this warp
  const int ox = oxw + lane * 4;
  const int oy0 = (blockIdx.y * (8 / wx) + warp / wx) * RS;
  if (oy0 >= out_h || oxw >= out_w) return;                           // whole warps only: the ring is per warp
  float kr[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) kr[i] = __ldg(k + 15 - i);           // flipped taps: kr[a*4+b] = k[3-a][3-b]
  // rank-1 test on the flipped taps: kr[a][b] == ky[a] * kx[b] with kx = row 0, ky = column 0 / kr[0][0]
  float kx[4], ky[4];
  bool separable = kr[0] != 0.f;
#pragma unroll
  for (int b = 0; b < 4; ++b) kx[b] = kr[b];
#pragma unroll
  for (int a = 0; a < 4; ++a) ky[a] = separable ? kr[a * 4] / kr[0] : 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
      separable = separable && fabsf(ky[a] * kx[b] - kr[a * 4 + b]) <= 1e-6f * fabsf(kr[a * 4 + b]) + 1e-30f;
  const float* xp = x + (size_t)blockIdx.z * in_h * in_w;
  float* yp = y + (size_t)blockIdx.z * out_h * out_w;
  const int ixw = oxw - px0;                                          // first input column of the warp's segment
  const bool vec_store = ((out_w & 3) == 0) && ox + 3 < out_w;
  const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(&ring[warp][0][0]);
  constexpr int NROWS = RS + 3;
  // row r of the strip (input row iy = oy0 - py0 + r) -> ring slot r & (kUpStages - 1).  The column part of every
  // copy (offset, in-bounds flag, shared address) is row-invariant and computed once
  int gxo[5];
  uint32_t dsto[5];
  bool okx[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const int e = j * 32 + lane, gx = ixw + e;
    okx[j] = e < 131 && gx >= 0 && gx < in_w;
    gxo[j] = okx[j] ? gx : 0;
    dsto[j] = ring_s + (uint32_t)(e < 131 ? e : 0) * 4u;
  }
  auto issue_row = [&](int r) {
    const int iy = oy0 - py0 + r;
    const bool row_ok = r < NROWS && iy >= 0 && iy < in_h;
    const float* rp = xp + (size_t)(row_ok ? iy : 0) * in_w;
    const uint32_t slot_off = (uint32_t)((r & (kUpStages - 1)) * kUpSeg) * 4u;
#pragma unroll
    for (int j = 0; j < 5; ++j)
      if (j < 4 || lane < 3)                                          // element 128 + lane exists for lanes 0..2 only
        cp_async4_zfill(dsto[j] + slot_off, rp + gxo[j], row_ok && okx[j]);
    cp_async_commit();
  };
#pragma unroll
  for (int r = 0; r < kUpStages - 1; ++r) issue_row(r);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 1
  for (int r0 = 0; r0 < NROWS; r0 += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u;
      issue_row(r + kUpStages - 1);                                   // always commits a (possibly empty) group
      cp_async_wait<kUpStages - 1>();                                 // row r has landed (this lane's copies) ...
      __syncwarp();                                                   // ... and every other lane's
      const float* seg = &ring[warp][r & (kUpStages - 1)][lane * 4];
      const float4 s0 = *reinterpret_cast<const float4*>(seg);
      const float4 s1 = *reinterpret_cast<const float4*>(seg + 4);
      const float v[7] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z};
      __syncwarp();                                                   // slot may be refilled by the next issue_row
      // input row r feeds output rows r - a (a = tap row); (r - a) & 3 == (u - a) & 3 because r0 % 4 == 0
      if (separable) {
        // rank-1 FIR (every make_kernel() blur: outer(k1, k1)): one horizontal pass per input row, then 4 scaled adds
        float hrow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float t = v[j] * kx[0];
#pragma unroll
          for (int b = 1; b < 4; ++b) t = fmaf(v[j + b], kx[b], t);
          hrow[j] = t;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[(u - a) & 3][j] = fmaf(hrow[j], ky[a], acc[(u - a) & 3][j]);
      } else {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[(u - a) & 3][j] = fmaf(v[j + b], kr[a * 4 + b], acc[(u - a) & 3][j]);
      }
      const int orow = r - 3;                                        // complete once its a = 3 row has arrived
      const int slot = (u + 1) & 3;
      if (orow >= 0 && orow < RS && oy0 + orow < out_h && ox < out_w) {
        float* dst = yp + (size_t)(oy0 + orow) * out_w + ox;
        if (vec_store) {
          __stcs(reinterpret_cast<float4*>(dst), make_float4(acc[slot][0], acc[slot][1], acc[slot][2], acc[slot][3]));
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (ox + j < out_w) dst[j] = acc[slot][j];
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[slot][j] = 0.f;
    }
  }
  cp_async_wait<0>();
}

// up = 1, down in {1, 2}, 4x4 taps, pads 0..3, planes up to 1040 columns wide: the input side done by the TMA engine.
// A strip of full-width rows of one plane is ONE contiguous span of global memory, so although its rows are only
// 4-byte aligned (width 2R+1 after the transposed convolution) the span can be fetched with 1-D bulk copies
// (`cp.async.bulk`, 16-byte aligned start and size: the span is widened to the enclosing 16-byte boundaries, which
// stay inside the tensor because the launch checks base alignment and total size).  One elected thread keeps up to
// S chunks of CR rows in flight per CTA (mbarrier complete_tx); there are no per-element copy instructions and no
// per-element address arithmetic -- the cp.async ring above spends 36 instructions per output on that, this kernel
// ~14 -- and the small stages (3 x 16 KB at 1025 columns) leave room for 4 CTAs per SM, which is what the HBM rate
// turned out to depend on (measured on the [256,1025,1025] blur: 8-row chunks x 3, 2 CTAs/SM 5.0 TB/s; 4-row chunks
// x 3, 4 CTAs/SM 6.1 TB/s = 0.93 of the copy peak; ring kernel 4.45).  A thread owns the output columns t, t+256, ...
// (conflict-free 4-byte LDS whatever the row's alignment; 2-way for down = 2) and walks down the strip with the
// separable (or general) FIR in registers exactly like the ring kernels; padding columns are handled by a masked path
// that only the threads owning an edge column take, padding rows contribute nothing and are never fetched.
__device__ __forceinline__ void bulk_load_1d(uint32_t dst_s, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_s), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// Rows [lo, hi) of the plane that chunk c of a strip holds (strip row r = input row iy0 + r, nrows rows, CR per chunk)
// and the 16-byte aligned global byte span [a0, a0 + bytes) that contains them.
struct BulkSpan {
  int lo, hi;
  size_t a0;
  uint32_t bytes;
};
__device__ __forceinline__ BulkSpan bulk_span(int c, int CR, int iy0, int nrows, int in_h, int in_w, size_t plane_off) {
  BulkSpan sp;
  sp.lo = max(iy0 + c * CR, 0);
  sp.hi = min(iy0 + min((c + 1) * CR, nrows), in_h);
  const size_t g0 = (plane_off + (size_t)sp.lo * in_w) * 4, g1 = (plane_off + (size_t)max(sp.hi, sp.lo) * in_w) * 4;
  sp.a0 = g0 & ~(size_t)15;
  sp.bytes = (uint32_t)(((g1 + 15) & ~(size_t)15) - sp.a0);
  return sp;
}

template <int DOWN, int NC, int RS, int CR, int S>            // RS output rows per CTA, CR input rows per chunk, S stages
__global__ void __launch_bounds__(256) upfirdn2d_k4_bulk_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                const float* __restrict__ k, int in_h, int in_w,
                                                                int out_h, int out_w, int px0, int py0,
                                                                uint32_t stage_bytes) {
  static_assert(CR % 4 == 0 && (DOWN == 1 || DOWN == 2), "chunk rows keep the accumulator rotation static");
  constexpr int HALO = DOWN == 1 ? 3 : 2;
  constexpr int NACC = DOWN == 1 ? 4 : 2;
  extern __shared__ __align__(128) unsigned char bulk_smem[];
  __shared__ uint64_t full[S];
  const int t = threadIdx.x;
  const int oy0 = blockIdx.y * RS;
  float kr[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) kr[i] = __ldg(k + 15 - i);           // flipped taps: kr[a*4+b] = k[3-a][3-b]
  float kx[4], ky[4];
  bool separable = kr[0] != 0.f;                                      // rank-1 test, as in the ring kernels
#pragma unroll
  for (int b = 0; b < 4; ++b) kx[b] = kr[b];
#pragma unroll
  for (int a = 0; a < 4; ++a) ky[a] = separable ? kr[a * 4] / kr[0] : 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
      separable = separable && fabsf(ky[a] * kx[b] - kr[a * 4 + b]) <= 1e-6f * fabsf(kr[a * 4 + b]) + 1e-30f;
  const size_t plane_off = (size_t)blockIdx.z * in_h * in_w;          // elements
  float* yp = y + (size_t)blockIdx.z * out_h * out_w;
  const int nrows = DOWN * min(RS, out_h - oy0) + HALO;               // strip rows that feed a stored output row
  const int nchunks = (nrows + CR - 1) / CR;
  const int iy0 = DOWN * oy0 - py0;                                   // input row of strip row 0 (>= -3)
  const uint32_t smem_s = smem_u32(bulk_smem);
  auto issue_chunk = [&](int c) {                                     // one thread
    const BulkSpan sp = bulk_span(c, CR, iy0, nrows, in_h, in_w, plane_off);
    if (sp.lo >= sp.hi) return;                                       // only trailing chunks can be empty
    const int s = c % S;
    mbar_expect_tx(&full[s], sp.bytes);
    bulk_load_1d(smem_s + (uint32_t)s * stage_bytes, reinterpret_cast<const unsigned char*>(x) + sp.a0, sp.bytes,
                 &full[s]);
  };
  if (t == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (t == 0)
    for (int c = 0; c < S && c < nchunks; ++c) issue_chunk(c);

  // column state (row-invariant): byte offset of tap 0 inside a row; edge columns take the masked path
  uint32_t off[NC];
  uint32_t edge = 0, live = 0;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int ox = t + 256 * j, ix = DOWN * ox - px0;
    const bool in_img = ox < out_w;
    const bool inside = ix >= 0 && ix + 3 < in_w;
    off[j] = (in_img && inside) ? (uint32_t)ix * 4u : 0u;
    if (in_img) live |= 1u << j;
    if (in_img && !inside) edge |= 1u << j;
  }
  const bool any_edge = edge != 0;
  float* dst_row = yp + (size_t)oy0 * out_w + t;                       // next output row to complete (they complete in order)
  const uint32_t row_bytes = (uint32_t)in_w * 4u;

  // The whole strip loop is instantiated once per FIR kind so that the separable test is not re-evaluated per row.
  auto run = [&](auto sep_tag) {
    constexpr bool SEP = decltype(sep_tag)::value;
    float acc[NACC][NC];
#pragma unroll
    for (int i = 0; i < NACC; ++i)
#pragma unroll
      for (int j = 0; j < NC; ++j) acc[i][j] = 0.f;
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
      const BulkSpan sp = bulk_span(c, CR, iy0, nrows, in_h, in_w, plane_off);
      const int s = c % S;
      if (sp.lo < sp.hi) mbar_wait(&full[s], (uint32_t)(c / S) & 1u);
      // shared address of strip row c*CR (may lie before the span when that row is padding: never dereferenced then)
      uint32_t row_s = smem_s + (uint32_t)s * stage_bytes +
                       (uint32_t)(int32_t)((int64_t)((plane_off + (size_t)sp.lo * in_w) * 4 - sp.a0) +
                                           (int64_t)(iy0 + c * CR - sp.lo) * (int64_t)row_bytes);
#pragma unroll
      for (int u8 = 0; u8 < CR; ++u8, row_s += row_bytes) {
        const int r = c * CR + u8;
        if (r < nrows) {                                               // block-uniform
          const int iy = iy0 + r;
          const bool row_ok = iy >= sp.lo && iy < sp.hi;               // block-uniform: padding rows add nothing
          // DOWN 1: input row r feeds output rows r - a (tap row a), slot (r - a) & 3.  DOWN 2: r = 2m + e feeds
          // output m (tap row e, slot cur) and m - 1 (tap row e + 2, slot prev).  Static: chunks start at multiples of 4.
          const int u = u8 & 3, e = u8 & 1, cur = (u8 >> 1) & 1, prev = cur ^ 1;
          if (row_ok) {
            float v[NC][4];
#pragma unroll
            for (int j = 0; j < NC; ++j)
#pragma unroll
              for (int b = 0; b < 4; ++b) v[j][b] = lds_f32(row_s + off[j] + 4u * b);
            if (any_edge) {                                            // the few threads that own an edge column redo it
#pragma unroll
              for (int j = 0; j < NC; ++j)
                if (edge & (1u << j)) {
                  const int ix = DOWN * (t + 256 * j) - px0;
#pragma unroll
                  for (int b = 0; b < 4; ++b)
                    v[j][b] = (ix + b >= 0 && ix + b < in_w) ? lds_f32(row_s + (uint32_t)(ix + b) * 4u) : 0.f;
                }
            }
            if constexpr (SEP) {
#pragma unroll
              for (int j = 0; j < NC; ++j) {
                float h = v[j][0] * kx[0];
#pragma unroll
                for (int b = 1; b < 4; ++b) h = fmaf(v[j][b], kx[b], h);
                if constexpr (DOWN == 1) {
#pragma unroll
                  for (int a = 0; a < 4; ++a) acc[(u - a) & 3][j] = fmaf(h, ky[a], acc[(u - a) & 3][j]);
                } else {
                  acc[cur][j] = fmaf(h, ky[e], acc[cur][j]);
                  acc[prev][j] = fmaf(h, ky[e + 2], acc[prev][j]);
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < NC; ++j)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                  if constexpr (DOWN == 1) {
#pragma unroll
                    for (int a = 0; a < 4; ++a)
                      acc[(u - a) & 3][j] = fmaf(v[j][b], kr[a * 4 + b], acc[(u - a) & 3][j]);
                  } else {
                    acc[cur][j] = fmaf(v[j][b], kr[e * 4 + b], acc[cur][j]);
                    acc[prev][j] = fmaf(v[j][b], kr[(e + 2) * 4 + b], acc[prev][j]);
                  }
                }
            }
          }
          // the output row that is complete once this input row has been added (row counts bound it by nrows)
          const int orow = DOWN == 1 ? r - 3 : (r >> 1) - 1;
          const int slot = DOWN == 1 ? (u + 1) & 3 : prev;
          if (DOWN == 1 || e == 1) {
            if (orow >= 0) {
#pragma unroll
              for (int j = 0; j < NC; ++j)
                if (live & (1u << j)) __stcs(dst_row + 256 * j, acc[slot][j]);
              dst_row += out_w;
            }
#pragma unroll
            for (int j = 0; j < NC; ++j) acc[slot][j] = 0.f;
          }
        }
      }
      __syncthreads();                                                 // every thread is done with stage s
      if (t == 0 && c + S < nchunks) issue_chunk(c + S);
    }
  };
  if (separable) run(std::true_type{});
  else run(std::false_type{});
}

// up = 2, down = 1, 4x4 taps, pad (2,1) (the RGB-skip Upsample) with the same TMA-fed strips: a thread owns the input
// columns t, t+256, ... and slides a 3-row x 3-column window; every input pixel yields a 2x2 output quad, stored as
// two 8-byte pairs (a warp writes 256 contiguous bytes per store).  Chunks are a multiple of 3 rows so that the
// window rotation is static.
template <int NC, int RS, int CR, int S>                      // RS INPUT rows per CTA
__global__ void __launch_bounds__(256) upfirdn2d_up2_k4_bulk_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                    const float* __restrict__ k, int in_h, int in_w,
                                                                    uint32_t stage_bytes) {
  static_assert(CR % 3 == 0, "chunk rows keep the window rotation static");
  extern __shared__ __align__(128) unsigned char bulk_smem[];
  __shared__ uint64_t full[S];
  const int t = threadIdx.x;
  const int y0 = blockIdx.y * RS;
  float kr[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) kr[i] = __ldg(k + 15 - i);
  const size_t plane_off = (size_t)blockIdx.z * in_h * in_w;
  const int out_w = 2 * in_w;
  float* yp = y + (size_t)blockIdx.z * (2 * in_h) * out_w;
  const int nrows = min(RS, in_h - y0) + 2;                           // strip row r = input row y0 - 1 + r
  const int nchunks = (nrows + CR - 1) / CR;
  const int iy0 = y0 - 1;
  const uint32_t smem_s = smem_u32(bulk_smem);
  auto issue_chunk = [&](int c) {
    const BulkSpan sp = bulk_span(c, CR, iy0, nrows, in_h, in_w, plane_off);
    if (sp.lo >= sp.hi) return;
    const int s = c % S;
    mbar_expect_tx(&full[s], sp.bytes);
    bulk_load_1d(smem_s + (uint32_t)s * stage_bytes, reinterpret_cast<const unsigned char*>(x) + sp.a0, sp.bytes,
                 &full[s]);
  };
  if (t == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (t == 0)
    for (int c = 0; c < S && c < nchunks; ++c) issue_chunk(c);
  uint32_t off[NC];
  uint32_t edge = 0, live = 0;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int xx = t + 256 * j;
    const bool in_img = xx < in_w;
    const bool inside = xx >= 1 && xx + 1 < in_w;
    off[j] = (in_img && inside) ? (uint32_t)(xx - 1) * 4u : 0u;
    if (in_img) live |= 1u << j;
    if (in_img && !inside) edge |= 1u << j;
  }
  const bool any_edge = edge != 0;
  const uint32_t row_bytes = (uint32_t)in_w * 4u;
  float* dst_row = yp + (size_t)(2 * y0) * out_w + 2 * t;             // output rows are produced in order
  float win[3][NC][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < NC; ++j)
#pragma unroll
      for (int b = 0; b < 3; ++b) win[i][j][b] = 0.f;
#pragma unroll 1
  for (int c = 0; c < nchunks; ++c) {
    const BulkSpan sp = bulk_span(c, CR, iy0, nrows, in_h, in_w, plane_off);
    const int s = c % S;
    if (sp.lo < sp.hi) mbar_wait(&full[s], (uint32_t)(c / S) & 1u);
    uint32_t row_s = smem_s + (uint32_t)s * stage_bytes +
                     (uint32_t)(int32_t)((int64_t)((plane_off + (size_t)sp.lo * in_w) * 4 - sp.a0) +
                                         (int64_t)(iy0 + c * CR - sp.lo) * (int64_t)row_bytes);
#pragma unroll
    for (int u8 = 0; u8 < CR; ++u8, row_s += row_bytes) {
      const int r = c * CR + u8;
      if (r < nrows) {
        const int u = u8 % 3;                                          // strip row r lives in window slot r % 3 = u
        const int iy = iy0 + r;
        const bool row_ok = iy >= sp.lo && iy < sp.hi;
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < NC; ++j)
#pragma unroll
            for (int b = 0; b < 3; ++b) win[u][j][b] = lds_f32(row_s + off[j] + 4u * b);
          if (any_edge) {
#pragma unroll
            for (int j = 0; j < NC; ++j)
              if (edge & (1u << j)) {
                const int ix = t + 256 * j - 1;
#pragma unroll
                for (int b = 0; b < 3; ++b)
                  win[u][j][b] = (ix + b >= 0 && ix + b < in_w) ? lds_f32(row_s + (uint32_t)(ix + b) * 4u) : 0.f;
              }
          }
        } else {
#pragma unroll
          for (int j = 0; j < NC; ++j)
#pragma unroll
            for (int b = 0; b < 3; ++b) win[u][j][b] = 0.f;
        }
        if (r >= 2) {                                                  // rows yy-1, yy, yy+1 = strip rows r-2, r-1, r
#pragma unroll
          for (int py = 0; py < 2; ++py) {
#pragma unroll
            for (int j = 0; j < NC; ++j) {
              float o[2];
#pragma unroll
              for (int px = 0; px < 2; ++px) {
                float sum = 0.f;
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                  for (int b = 0; b < 2; ++b)               // strip row r - 2 + py + a -> slot (u + 1 + py + a) % 3
                    sum = fmaf(win[(u + 1 + py + a) % 3][j][px + b], kr[(py + 2 * a) * 4 + px + 2 * b], sum);
                o[px] = sum;
              }
              if (live & (1u << j)) __stcs(reinterpret_cast<float2*>(dst_row + 512 * j), make_float2(o[0], o[1]));
            }
            dst_row += out_w;
          }
        }
      }
    }
    __syncthreads();
    if (t == 0 && c + S < nchunks) issue_chunk(c + S);
  }
}

// up = 1, down = 2, 4x4 taps, pad >= 0 (the Downsample / ConvLayer blur of the reference, model.py:56-74; not on the
// swap path): same scheme as the up1 kernel -- per-warp cp.async ring, register accumulators -- with two input rows
// and 2 x 128 + 2 input columns per output row.  Input row r = 2m + e feeds output row m with tap row e and output
// row m - 1 with tap row e + 2, so two accumulator rows are live.
constexpr int kDnStages = 4;
constexpr int kDnSeg = 264;                                  // 258 floats per row segment, padded

template <int RS>
__global__ void __launch_bounds__(256) upfirdn2d_down2_k4_strip_kernel(const float* __restrict__ x,
                                                                       float* __restrict__ y,
                                                                       const float* __restrict__ k, int in_h, int in_w,
                                                                       int out_h, int out_w, int px0, int py0) {
  __shared__ __align__(16) float ring[8][kDnStages][kDnSeg];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int oxw = blockIdx.x * 128;
  const int ox = oxw + lane * 4;
  const int oy0 = (blockIdx.y * 8 + warp) * RS;
  if (oy0 >= out_h) return;                                           // whole warps only: the ring is per warp
  float kr[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) kr[i] = __ldg(k + 15 - i);           // flipped taps: kr[a*4+b] = k[3-a][3-b]
  float kx[4], ky[4];
  bool separable = kr[0] != 0.f;
#pragma unroll
  for (int b = 0; b < 4; ++b) kx[b] = kr[b];
#pragma unroll
  for (int a = 0; a < 4; ++a) ky[a] = separable ? kr[a * 4] / kr[0] : 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
      separable = separable && fabsf(ky[a] * kx[b] - kr[a * 4 + b]) <= 1e-6f * fabsf(kr[a * 4 + b]) + 1e-30f;
  const float* xp = x + (size_t)blockIdx.z * in_h * in_w;
  float* yp = y + (size_t)blockIdx.z * out_h * out_w;
  const int ixw = 2 * oxw - px0;                                      // first input column of the warp's segment
  const int iy_first = 2 * oy0 - py0;                                 // input row of strip row 0
  const bool vec_store = ((out_w & 3) == 0) && ox + 3 < out_w;
  const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(&ring[warp][0][0]);
  constexpr int NROWS = 2 * RS + 2;
  int gxo[9];
  bool okx[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    const int e = j * 32 + lane, gx = ixw + e;
    okx[j] = e < 258 && gx >= 0 && gx < in_w;
    gxo[j] = okx[j] ? gx : 0;
  }
  auto issue_row = [&](int r) {
    const int iy = iy_first + r;
    const bool row_ok = r < NROWS && iy >= 0 && iy < in_h;
    const float* rp = xp + (size_t)(row_ok ? iy : 0) * in_w;
    const uint32_t dst = ring_s + (uint32_t)((r & (kDnStages - 1)) * kDnSeg + lane) * 4u;
#pragma unroll
    for (int j = 0; j < 9; ++j)
      if (j < 8 || lane < 2)                                          // elements 256, 257 exist for lanes 0, 1 only
        cp_async4_zfill(dst + (uint32_t)(j * 32) * 4u, rp + gxo[j], row_ok && okx[j]);
    cp_async_commit();
  };
#pragma unroll
  for (int r = 0; r < kDnStages - 1; ++r) issue_row(r);
  float acc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 1
  for (int r0 = 0; r0 < NROWS; r0 += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u;                                           // r = 2m + e with m = r0/2 + (u >> 1), e = u & 1
      issue_row(r + kDnStages - 1);
      cp_async_wait<kDnStages - 1>();
      __syncwarp();
      const float* seg = &ring[warp][r & (kDnStages - 1)][lane * 8];
      const float4 s0 = *reinterpret_cast<const float4*>(seg);
      const float4 s1 = *reinterpret_cast<const float4*>(seg + 4);
      const float4 s2 = *reinterpret_cast<const float4*>(seg + 8);
      const float v[10] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w, s2.x, s2.y};
      __syncwarp();
      const int e = u & 1, cur = (u >> 1) & 1, prev = cur ^ 1;       // output m -> slot cur, output m - 1 -> slot prev
      if (separable) {
        float hrow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float t = v[2 * j] * kx[0];
#pragma unroll
          for (int b = 1; b < 4; ++b) t = fmaf(v[2 * j + b], kx[b], t);
          hrow[j] = t;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[cur][j] = fmaf(hrow[j], ky[e], acc[cur][j]);
          acc[prev][j] = fmaf(hrow[j], ky[e + 2], acc[prev][j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            acc[cur][j] = fmaf(v[2 * j + b], kr[e * 4 + b], acc[cur][j]);
            acc[prev][j] = fmaf(v[2 * j + b], kr[(e + 2) * 4 + b], acc[prev][j]);
          }
      }
      if (e == 1) {                                                   // rows 2m-2 .. 2m+1 seen: output m - 1 is complete
        const int orow = (r >> 1) - 1;
        if (orow >= 0 && orow < RS && oy0 + orow < out_h && ox < out_w) {
          float* dst = yp + (size_t)(oy0 + orow) * out_w + ox;
          if (vec_store) {
            __stcs(reinterpret_cast<float4*>(dst), make_float4(acc[prev][0], acc[prev][1], acc[prev][2], acc[prev][3]));
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (ox + j < out_w) dst[j] = acc[prev][j];
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[prev][j] = 0.f;
      }
    }
  }
  cp_async_wait<0>();
}

// up = 2, down = 1, 4x4 taps, pad (2,1): the RGB-skip Upsample (model.py:35-53), out = 2 x in.  Output parity (py,px)
// has 2x2 live taps: out[2y+py][2x+px] = sum_{a,b in {0,1}} in[y-1+py+a][x-1+px+b] * kf[py+2a][px+2b].
// A lane owns 4 input columns (8 output columns, two float4 stores per output row) and slides a 3-row window.
template <int RS>
__global__ void __launch_bounds__(256) upfirdn2d_up2_k4_strip_kernel(const float* __restrict__ x,
                                                                     float* __restrict__ y,
                                                                     const float* __restrict__ k, int in_h, int in_w,
                                                                     int wx) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = (blockIdx.x * wx + (warp & (wx - 1))) * 128 + lane * 4;   // wx warps side by side (see the up1 kernel)
  const int y0 = (blockIdx.y * (8 / wx) + warp / wx) * RS;
  if (y0 >= in_h || x0 >= in_w) return;
  float kr[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) kr[i] = __ldg(k + 15 - i);
  const float* xp = x + (size_t)blockIdx.z * in_h * in_w;
  const int out_w = 2 * in_w;
  float* yp = y + (size_t)blockIdx.z * (2 * in_h) * out_w;
  const bool interior = x0 >= 1 && x0 + 4 < in_w;
  const bool full = x0 + 3 < in_w && (in_w & 1) == 0;               // 2*in_w % 4 == 0 and all 8 outputs exist
  float win[3][6];
  auto load_row = [&](int iy, float* v) {
    const bool row_ok = iy >= 0 && iy < in_h;
    const float* rp = xp + (size_t)(row_ok ? iy : 0) * in_w + x0 - 1;
    if (row_ok && interior) {
#pragma unroll
      for (int c = 0; c < 6; ++c) v[c] = __ldg(rp + c);
    } else {
#pragma unroll
      for (int c = 0; c < 6; ++c) v[c] = (row_ok && x0 - 1 + c >= 0 && x0 - 1 + c < in_w) ? __ldg(rp + c) : 0.f;
    }
  };
  load_row(y0 - 1, win[0]);
  load_row(y0, win[1]);
#pragma unroll 1
  for (int r0 = 0; r0 < RS; r0 += 3) {
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int yy = y0 + r0 + u;
      // rows yy-1, yy, yy+1 live in slots u % 3, (u+1) % 3, (u+2) % 3
      load_row((r0 + u < RS && yy < in_h) ? yy + 1 : -1, win[(u + 2) % 3]);
      if (r0 + u < RS && yy < in_h) {
#pragma unroll
        for (int py = 0; py < 2; ++py) {
          float o[8];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int px = 0; px < 2; ++px) {
              float s = 0.f;
#pragma unroll
              for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b)
                  s = fmaf(win[(u + py + a) % 3][i + px + b], kr[(py + 2 * a) * 4 + px + 2 * b], s);
              o[2 * i + px] = s;
            }
          float* dst = yp + (size_t)(2 * yy + py) * out_w + 2 * x0;
          if (full) {
            __stcs(reinterpret_cast<float4*>(dst), make_float4(o[0], o[1], o[2], o[3]));
            __stcs(reinterpret_cast<float4*>(dst) + 1, make_float4(o[4], o[5], o[6], o[7]));
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (2 * x0 + j < out_w) dst[j] = o[j];
          }
        }
      }
    }
  }
}

// General path: any up/down/pad (incl. negative pads = crop) and any kernel size.
__global__ void __launch_bounds__(256) upfirdn2d_general_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                const float* __restrict__ k, int in_h, int in_w,
                                                                int out_h, int out_w, int kh, int kw, int up_x,
                                                                int up_y, int down_x, int down_y, int px0, int py0,
                                                                int64_t total) {
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int ox = (int)(idx % out_w);
    int64_t t = idx / out_w;
    int oy = (int)(t % out_h);
    int plane = (int)(t / out_h);
    const float* xp = x + (size_t)plane * in_h * in_w;
    float acc = 0.f;
    for (int ky = 0; ky < kh; ++ky) {
      int uy = oy * down_y + ky - py0;          // coordinate in the zero-stuffed image
      if (uy < 0 || uy % up_y != 0) continue;
      int iy = uy / up_y;
      if (iy >= in_h) continue;
      for (int kx = 0; kx < kw; ++kx) {
        int ux = ox * down_x + kx - px0;
        if (ux < 0 || ux % up_x != 0) continue;
        int ix = ux / up_x;
        if (ix >= in_w) continue;
        acc = fmaf(__ldg(xp + (size_t)iy * in_w + ix), __ldg(k + (kh - 1 - ky) * kw + (kw - 1 - kx)), acc);
      }
    }
    y[idx] = acc;
  }
}

int launch_upfirdn2d(const float* x, float* y, const float* k, int planes, int in_h, int in_w, int kh, int kw,
                     int up_x, int up_y, int down_x, int down_y, int px0, int px1, int py0, int py1,
                     cudaStream_t st) {
  if (planes == 0) return HF_OK;
  HF_REQUIRE(x && y && k, "upfirdn2d: null pointer");
  HF_REQUIRE(planes >= 0 && in_h > 0 && in_w > 0 && kh > 0 && kw > 0, "upfirdn2d: bad shape");
  HF_REQUIRE(up_x > 0 && up_y > 0 && down_x > 0 && down_y > 0, "upfirdn2d: up/down must be positive");
  const int out_h = (in_h * up_y + py0 + py1 - kh) / down_y + 1;
  const int out_w = (in_w * up_x + px0 + px1 - kw) / down_x + 1;
  HF_REQUIRE(in_h * up_y + py0 + py1 - kh >= 0 && in_w * up_x + px0 + px1 - kw >= 0,
             "upfirdn2d: output would be empty (in %dx%d up %d pad %d,%d k %d)", in_h, in_w, up_y, py0, py1, kh);
  if (planes == 0) return HF_OK;
  const bool k4 = (kh == 4 && kw == 4);
  const bool sym = (up_x == up_y && down_x == down_y);
  const bool pads03 = px0 >= 0 && py0 >= 0 && px1 >= 0 && py1 >= 0 && px0 <= 3 && px1 <= 3 && py0 <= 3 && py1 <= 3;
  // TMA-fed strip kernels (see upfirdn2d_k4_bulk_kernel): full-width rows, 16-byte aligned tensor, pads 0..3
  const bool bulk_ok = k4 && sym && planes <= 65535 && in_w >= 64 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                       ((size_t)planes * in_h * in_w) % 4 == 0;
  auto bulk_stage = [&](int CR) {
    return (uint32_t)((((size_t)CR * in_w * 4 + 15) / 16 * 16 + 32 + 127) / 128 * 128);
  };
  auto bulk_go = [&](auto kern, int rows_per_cta, int CR, int S, auto... tail) {
    const uint32_t stage = bulk_stage(CR);                 // <= 25 KB for the widths admitted below
    // per launch like the convolution kernels: the attribute is per device, a cache would have to be too
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    dim3 grid(1, cdiv(up_x == 2 ? in_h : out_h, rows_per_cta), planes);
    kern<<<grid, 256, S * stage, st>>>(x, y, k, in_h, in_w, tail..., stage);
  };
  if (bulk_ok && up_x == 1 && down_x == 1 && pads03 && out_w <= 1024 && in_w <= 1040) {
    if (out_w <= 512) bulk_go(upfirdn2d_k4_bulk_kernel<1, 2, 64, 4, 3>, 64, 4, 3, out_h, out_w, px0, py0);
    else bulk_go(upfirdn2d_k4_bulk_kernel<1, 4, 64, 4, 3>, 64, 4, 3, out_h, out_w, px0, py0);
  } else if (bulk_ok && up_x == 1 && down_x == 2 && pads03 && out_w <= 1024 && in_w <= 1040) {
    if (out_w <= 512) bulk_go(upfirdn2d_k4_bulk_kernel<2, 2, 32, 4, 3>, 32, 4, 3, out_h, out_w, px0, py0);
    else bulk_go(upfirdn2d_k4_bulk_kernel<2, 4, 32, 4, 3>, 32, 4, 3, out_h, out_w, px0, py0);
  } else if (bulk_ok && up_x == 2 && down_x == 1 && px0 == 2 && py0 == 2 && px1 == 1 && py1 == 1 && in_w <= 1024 &&
             (reinterpret_cast<uintptr_t>(y) & 7) == 0) {
    if (in_w <= 512) bulk_go(upfirdn2d_up2_k4_bulk_kernel<2, 32, 6, 3>, 32, 6, 3);
    else bulk_go(upfirdn2d_up2_k4_bulk_kernel<4, 32, 6, 3>, 32, 6, 3);
  } else if (k4 && sym && up_x == 1 && down_x == 1 && planes <= 65535 && px0 >= 0 && py0 >= 0 && px1 >= 0 && py1 >= 0) {
    constexpr int RS = 64;                                   // rows per warp strip (3 halo rows each)
    const int wx = out_w > 512 ? 8 : out_w > 256 ? 4 : out_w > 128 ? 2 : 1;   // warps side by side
    dim3 grid(cdiv(out_w, 128 * wx), cdiv(out_h, RS * (8 / wx)), planes);
    upfirdn2d_up1_k4_strip_kernel<RS><<<grid, 256, 0, st>>>(x, y, k, in_h, in_w, out_h, out_w, px0, py0, wx);
  } else if (k4 && sym && up_x == 2 && down_x == 1 && planes <= 65535 && px0 == 2 && py0 == 2 && px1 == 1 &&
             py1 == 1) {
    constexpr int RS = 12;                                   // input rows per warp strip
    const int wx = 1;              // warps stacked vertically (measured: 4.9 TB/s vs 4.5 side by side for 512-wide planes)
    dim3 grid(cdiv(in_w, 128 * wx), cdiv(in_h, RS * (8 / wx)), planes);
    upfirdn2d_up2_k4_strip_kernel<RS><<<grid, 256, 0, st>>>(x, y, k, in_h, in_w, wx);
  } else if (k4 && sym && up_x == 1 && down_x == 2 && planes <= 65535 && px0 >= 0 && py0 >= 0 && px1 >= 0 && py1 >= 0) {
    constexpr int RS = 16;                                   // 8 warps x 16 output rows x 128 output columns per CTA
    dim3 grid(cdiv(out_w, 128), cdiv(out_h, 8 * RS), planes);
    upfirdn2d_down2_k4_strip_kernel<RS><<<grid, 256, 0, st>>>(x, y, k, in_h, in_w, out_h, out_w, px0, py0);
  } else if (k4 && sym && up_x == 1 && (down_x == 1 || down_x == 2) && planes <= 65535) {
    if (down_x == 1) {
      dim3 grid(cdiv(out_w, 64), cdiv(out_h, 32), planes);
      upfirdn2d_up1_k4_kernel<1, 32><<<grid, 256, 0, st>>>(x, y, k, in_h, in_w, out_h, out_w, px0, py0);
    } else {
      dim3 grid(cdiv(out_w, 64), cdiv(out_h, 16), planes);
      upfirdn2d_up1_k4_kernel<2, 16><<<grid, 256, 0, st>>>(x, y, k, in_h, in_w, out_h, out_w, px0, py0);
    }
  } else if (k4 && sym && up_x == 2 && down_x == 1 &&
             (int64_t)planes * out_h * ((out_w + 3) / 4) < (int64_t)2000000000) {
    int64_t quads = (int64_t)planes * out_h * ((out_w + 3) / 4);
    int grid = (int)std::min<int64_t>((quads + 255) / 256, (int64_t)num_sms() * 32);
    upfirdn2d_up2_k4_kernel<<<grid, 256, 0, st>>>(x, y, k, in_h, in_w, out_h, out_w, px0, py0, quads);
  } else {
    int64_t total = (int64_t)planes * out_h * out_w;
    int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 16);
    upfirdn2d_general_kernel<<<grid, 256, 0, st>>>(x, y, k, in_h, in_w, out_h, out_w, kh, kw, up_x, up_y,
                                                    down_x, down_y, px0, py0, total);
  }
  HF_LAUNCH_OK("upfirdn2d");
  count_launch();
  return HF_OK;
}

// =============================================================================================
// fused bias + activation  (op/fused_bias_act_kernel.cu:19-49, forward, act in {1 linear, 3 lrelu})
// =============================================================================================
template <bool VEC>
__global__ void __launch_bounds__(256) bias_act_kernel(const float* __restrict__ x, const float* __restrict__ b,
                                                       float* __restrict__ y, int64_t n, int size_b,
                                                       int64_t step_b, int act, float alpha, float scale) {
  if (VEC) {
    const int64_t n4 = n >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
      float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
      float bb = size_b ? __ldg(b + ((i * 4) / step_b) % size_b) : 0.f;   // step_b % 4 == 0: one channel
      float r[4] = {v.x + bb, v.y + bb, v.z + bb, v.w + bb};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t = r[j];
        if (act == 3) t = (t > 0.f) ? t : t * alpha;
        r[j] = t * scale;
      }
      reinterpret_cast<float4*>(y)[i] = make_float4(r[0], r[1], r[2], r[3]);
    }
  } else {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      float t = x[i];
      if (size_b) t += __ldg(b + (i / step_b) % size_b);
      if (act == 3) t = (t > 0.f) ? t : t * alpha;
      y[i] = t * scale;
    }
  }
}

// Plane form of the vector path: blockIdx.y = (batch, channel) plane, so the bias index is computed once per CTA
// instead of two 64-bit divisions per float4 (measured: the flat form ran at about a third of the HBM rate).
__global__ void __launch_bounds__(256) bias_act_plane_kernel(const float* __restrict__ x, const float* __restrict__ b,
                                                             float* __restrict__ y, int quads_per_plane, int size_b,
                                                             int act, float alpha, float scale) {
  const int plane = blockIdx.y;
  const float bb = size_b ? __ldg(b + plane % size_b) : 0.f;
  const float4* xp = reinterpret_cast<const float4*>(x) + (size_t)plane * quads_per_plane;
  float4* yp = reinterpret_cast<float4*>(y) + (size_t)plane * quads_per_plane;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < quads_per_plane; i += gridDim.x * blockDim.x) {
    const float4 v = __ldg(xp + i);
    float r[4] = {v.x + bb, v.y + bb, v.z + bb, v.w + bb};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float t = r[j];
      if (act == 3) t = (t > 0.f) ? t : t * alpha;
      r[j] = t * scale;
    }
    yp[i] = make_float4(r[0], r[1], r[2], r[3]);
  }
}

int launch_bias_act(const float* x, const float* b, float* y, int64_t n, int size_b, int64_t step_b, int act,
                    float alpha, float scale, cudaStream_t st) {
  if (n == 0) return HF_OK;
  if (x && y && (size_b == 0 || b) && (act == 1 || act == 3) && step_b % 4 == 0 && step_b >= 1024 && n % step_b == 0 &&
      n / step_b <= 65535 && step_b / 4 < (int64_t)2000000000 && ((((uintptr_t)x | (uintptr_t)y) & 15) == 0)) {
    const int planes = (int)(n / step_b), quads = (int)(step_b / 4);
    const int gx = std::max(1, std::min(cdiv(quads, 256), std::max(1, num_sms() * 16 / planes)));
    bias_act_plane_kernel<<<dim3(gx, planes), 256, 0, st>>>(x, b, y, quads, size_b, act, alpha, scale);
    HF_LAUNCH_OK("bias_act");
    count_launch();
    return HF_OK;
  }
  HF_REQUIRE(x && y, "bias_act: null pointer");
  HF_REQUIRE(act == 1 || act == 3, "bias_act: act must be 1 (linear) or 3 (leaky relu), got %d", act);
  HF_REQUIRE(n >= 0 && size_b >= 0 && step_b >= 1, "bias_act: bad sizes");
  HF_REQUIRE(size_b == 0 || b, "bias_act: bias pointer is null");
  if (n == 0) return HF_OK;
  const bool vec = (step_b % 4 == 0) && (n % 4 == 0) && ((((uintptr_t)x | (uintptr_t)y) & 15) == 0);
  int64_t work = vec ? n / 4 : n;
  int grid = (int)std::min<int64_t>((work + 255) / 256, (int64_t)num_sms() * 16);
  if (vec)
    bias_act_kernel<true><<<grid, 256, 0, st>>>(x, b, y, n, size_b, step_b, act, alpha, scale);
  else
    bias_act_kernel<false><<<grid, 256, 0, st>>>(x, b, y, n, size_b, step_b, act, alpha, scale);
  HF_LAUNCH_OK("bias_act");
  count_launch();
  return HF_OK;
}

// =============================================================================================
// style affine (EqualLinear modulation, model.py:153-163 with bias_init=1) and demodulation table
// =============================================================================================
struct AffineJobs { AffineJob j[kMaxJobs]; int n; };
struct DemodJobs { DemodJob j[kMaxJobs]; int n; };

// One warp per output row i of one job; lanes split the D-long dot product (float4 loads).
__global__ void __launch_bounds__(256) affine_kernel(const __grid_constant__ AffineJobs jobs, int B, int D,
                                                     int64_t style_stride) {
  int jb = 0;
  while (jb + 1 < jobs.n && (int)blockIdx.x >= jobs.j[jb + 1].block_begin) ++jb;
  const AffineJob& job = jobs.j[jb];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = ((int)blockIdx.x - job.block_begin) * 8 + warp;
  if (i >= job.C) return;
  const float* wrow = job.mw + (size_t)i * D;
  const float mb = __ldg(job.mb + i);
  if (D == 512) {
    // the usual case (style_dim 512): the weight row lives in registers, four samples in flight per pass
    float4 w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) w[q] = __ldg(reinterpret_cast<const float4*>(wrow + lane * 4 + q * 128));
    for (int b0 = 0; b0 < B; b0 += 4) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int b = min(b0 + u, B - 1);
        const float* srow = job.style + (size_t)b * style_stride;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 s = __ldg(reinterpret_cast<const float4*>(srow + lane * 4 + q * 128));
          acc[u] = fmaf(w[q].x, s.x, acc[u]); acc[u] = fmaf(w[q].y, s.y, acc[u]);
          acc[u] = fmaf(w[q].z, s.z, acc[u]); acc[u] = fmaf(w[q].w, s.w, acc[u]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
      }
      if (lane < 4 && b0 + lane < B) {
        const float a = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
        job.s[(size_t)(b0 + lane) * job.C + i] = fmaf(a, job.wscale, mb);
      }
    }
    return;
  }
  for (int b = 0; b < B; ++b) {
    const float* srow = job.style + (size_t)b * style_stride;
    float acc = 0.f;
    for (int j = lane * 4; j < D; j += 128) {
      float4 w = __ldg(reinterpret_cast<const float4*>(wrow + j));
      float4 s = __ldg(reinterpret_cast<const float4*>(srow + j));
      acc = fmaf(w.x, s.x, acc); acc = fmaf(w.y, s.y, acc); acc = fmaf(w.z, s.z, acc); acc = fmaf(w.w, s.w, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) job.s[(size_t)b * job.C + i] = fmaf(acc, job.wscale, mb);
  }
}

__global__ void __launch_bounds__(256) demod_kernel(const __grid_constant__ DemodJobs jobs, int B) {
  int jb = 0;
  while (jb + 1 < jobs.n && (int)blockIdx.x >= jobs.j[jb + 1].block_begin) ++jb;
  const DemodJob& job = jobs.j[jb];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = ((int)blockIdx.x - job.block_begin) * 8 + warp;
  if (o >= job.Cout) return;
  const float* wrow = job.wsq + (size_t)o * job.Cin;
  if (job.Cin <= 512) {
    // Wsq row in registers (<= 16 values per lane), four samples in flight per pass; same summation order as below
    float w[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) w[q] = (lane + 32 * q < job.Cin) ? __ldg(wrow + lane + 32 * q) : 0.f;
    for (int b0 = 0; b0 < B; b0 += 4) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* srow = job.s + (size_t)min(b0 + u, B - 1) * job.Cin;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          if (lane + 32 * q < job.Cin) {
            const float s = __ldg(srow + lane + 32 * q);
            acc[u] = fmaf(s * s, w[q], acc[u]);
          }
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], off);
      }
      if (lane < 4 && b0 + lane < B) {
        const float a = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
        job.d[(size_t)(b0 + lane) * job.Cout + o] = rsqrtf(a + 1e-8f);
      }
    }
    return;
  }
  for (int b = 0; b < B; ++b) {
    const float* srow = job.s + (size_t)b * job.Cin;
    float acc = 0.f;
    for (int i = lane; i < job.Cin; i += 32) {
      float s = __ldg(srow + i);
      acc = fmaf(s * s, __ldg(wrow + i), acc);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) job.d[(size_t)b * job.Cout + o] = rsqrtf(acc + 1e-8f);
  }
}

int launch_affine(AffineJob* jobs, int njobs, int B, int D, int64_t style_stride, cudaStream_t st) {
  HF_REQUIRE(njobs > 0 && njobs <= kMaxJobs, "affine: bad job count %d", njobs);
  HF_REQUIRE(D % 4 == 0 && style_stride % 4 == 0, "affine: style_dim and stride must be multiples of 4");
  AffineJobs aj;
  aj.n = njobs;
  int blocks = 0;
  for (int i = 0; i < njobs; ++i) {
    jobs[i].block_begin = blocks;
    blocks += cdiv(jobs[i].C, 8);
    aj.j[i] = jobs[i];
  }
  affine_kernel<<<blocks, 256, 0, st>>>(aj, B, D, style_stride);
  HF_LAUNCH_OK("affine");
  count_launch();
  return HF_OK;
}

int launch_demod(DemodJob* jobs, int njobs, int B, cudaStream_t st) {
  HF_REQUIRE(njobs > 0 && njobs <= kMaxJobs, "demod: bad job count %d", njobs);
  DemodJobs dj;
  dj.n = njobs;
  int blocks = 0;
  for (int i = 0; i < njobs; ++i) {
    jobs[i].block_begin = blocks;
    blocks += cdiv(jobs[i].Cout, 8);
    dj.j[i] = jobs[i];
  }
  demod_kernel<<<blocks, 256, 0, st>>>(dj, B);
  HF_LAUNCH_OK("demod");
  count_launch();
  return HF_OK;
}

// =============================================================================================
// pre-pass: NCHW fp32 -> NHWC 16-bit with the per-(b,cin) style scale folded in
// (the "x * s" of y = d * conv(x*s, W~)); optional FSE feature blend.
// =============================================================================================
template <int DT>
__global__ void __launch_bounds__(256) modulate_to_nhwc_kernel(const float* __restrict__ x, int64_t x_bstride,
                                                               const float* __restrict__ s,
                                                               const float* __restrict__ feat, float alpha,
                                                               uint16_t* __restrict__ xh, int C, int HW) {
  __shared__ float tile[64][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 64, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* xb = x + (size_t)b * x_bstride;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int c = c0 + ty + 8 * k, p = p0 + tx;
    float v = 0.f;
    if (c < C && p < HW) {
      v = __ldg(xb + (size_t)c * HW + p);
      if (feat) v = (1.f - alpha) * v + alpha * __ldg(feat + ((size_t)b * C + c) * HW + p);
      v *= __ldg(s + (size_t)b * C + c);
    }
    tile[ty + 8 * k][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int p = p0 + ty + 8 * k, c = c0 + 2 * tx;
    if (p < HW && c < C) {
      uint32_t packed = Half2T<DT>::pack(tile[2 * tx][ty + 8 * k], tile[2 * tx + 1][ty + 8 * k]);
      *reinterpret_cast<uint32_t*>(xh + ((size_t)b * HW + p) * C + c) = packed;
    }
  }
}

int launch_modulate_to_nhwc(const float* x, int x_broadcast, const float* s, const float* feat, float alpha,
                            void* xh, int B, int C, int HW, int dtype, cudaStream_t st) {
  HF_REQUIRE(x && s && xh, "modulate_to_nhwc: null pointer");
  HF_REQUIRE(C % 2 == 0, "modulate_to_nhwc: C must be even");
  dim3 grid(cdiv(HW, 32), cdiv(C, 64), B);
  int64_t bstride = x_broadcast ? 0 : (int64_t)C * HW;
  if (dtype == HF_BF16)
    modulate_to_nhwc_kernel<HF_BF16><<<grid, 256, 0, st>>>(x, bstride, s, feat, alpha, (uint16_t*)xh, C, HW);
  else
    modulate_to_nhwc_kernel<HF_F16><<<grid, 256, 0, st>>>(x, bstride, s, feat, alpha, (uint16_t*)xh, C, HW);
  HF_LAUNCH_OK("modulate_to_nhwc");
  count_launch();
  return HF_OK;
}

// =============================================================================================
// RGB combine: bias + ordered sum of the per-N-tile ToRGB partials + Upsample(skip)
// (ToRGB.forward model.py:356-365; Upsample = upfirdn2d(up=2, pad=(2,1)), model.py:35-53)
// =============================================================================================
__global__ void __launch_bounds__(256) rgb_combine_kernel(const float* partial, int num_partials,
                                                          int64_t partial_stride, const float* __restrict__ bias,
                                                          const float* __restrict__ skip,
                                                          const float* __restrict__ upk, float* rgb,
                                                          int H, int W, int64_t total) {
  __shared__ float kf[16];
  if (threadIdx.x < 16) kf[threadIdx.x] = upk ? upk[15 - threadIdx.x] : 0.f;
  __syncthreads();
  const int h2 = H >> 1, w2 = W >> 1;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int X = (int)(idx % W);
    int64_t t = idx / W;
    int Y = (int)(t % H);
    int plane = (int)(t / H);            // b*3 + j
    float acc = bias ? __ldg(bias + plane % 3) : 0.f;
    for (int q = 0; q < num_partials; ++q) acc += partial[q * partial_stride + idx];   // may alias rgb (in place)
    if (skip) {
      const float* sp = skip + (size_t)plane * h2 * w2;
      const int ky0 = Y & 1, kx0 = X & 1;      // pad0 = 2: Y + ky - 2 even
      float up = 0.f;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        int ky = ky0 + 2 * a;
        int uy = Y + ky - 2;
        int iy = uy >> 1;
        if (uy < 0 || iy >= h2) continue;
#pragma unroll
        for (int bq = 0; bq < 2; ++bq) {
          int kx = kx0 + 2 * bq;
          int ux = X + kx - 2;
          int ix = ux >> 1;
          if (ux >= 0 && ix < w2) up = fmaf(__ldg(sp + (size_t)iy * w2 + ix), kf[ky * 4 + kx], up);
        }
      }
      acc += up;
    }
    rgb[idx] = acc;
  }
}

// W % 4 == 0: four consecutive outputs per thread, float4 partial loads / stores; the 2x2-tap up-FIR of the
// skip reads a 2x4 window of the half-resolution image (columns 2q-1 .. 2q+2 for outputs 4q .. 4q+3).
__global__ void __launch_bounds__(256) rgb_combine_vec4_kernel(const float* partial, int num_partials,
                                                               int64_t partial_stride,
                                                               const float* __restrict__ bias,
                                                               const float* __restrict__ skip,
                                                               const float* __restrict__ upk, float* rgb, int H, int W,
                                                               int64_t total4) {
  __shared__ float kf[16];
  if (threadIdx.x < 16) kf[threadIdx.x] = upk ? upk[15 - threadIdx.x] : 0.f;
  __syncthreads();
  const int h2 = H >> 1, w2 = W >> 1, wq = W >> 2;
  // 32-bit index math (the launcher checks total4 < 2^31): a 64-bit div/mod costs ~100 instructions
  for (int i4 = blockIdx.x * blockDim.x + threadIdx.x; i4 < (int)total4; i4 += gridDim.x * blockDim.x) {
    const int q = i4 % wq;
    const int t = i4 / wq;
    const int Y = t % H;
    const int plane = t / H;
    const float bv = bias ? __ldg(bias + plane % 3) : 0.f;
    float acc[4] = {bv, bv, bv, bv};
    for (int s = 0; s < num_partials; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(partial + s * partial_stride + (size_t)i4 * 4);   // may alias rgb
      acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
    }
    if (skip) {
      const float* sp = skip + (size_t)plane * h2 * w2;
      const int ky0 = Y & 1;
      float up[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int ky = ky0 + 2 * a;
        const int uy = Y + ky - 2;
        const int iy = uy >> 1;
        if (uy < 0 || iy >= h2) continue;
        const float* row = sp + (size_t)iy * w2;
        float c[4];                                      // columns 2q-1 .. 2q+2
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int ix = 2 * q - 1 + m;
          c[m] = (ix >= 0 && ix < w2) ? __ldg(row + ix) : 0.f;
        }
        // even X (4q, 4q+2): taps kx 0,2 on columns (X/2-1, X/2); odd X: taps kx 1,3 on ((X-1)/2, (X+1)/2)
        up[0] = fmaf(c[0], kf[ky * 4 + 0], fmaf(c[1], kf[ky * 4 + 2], up[0]));
        up[1] = fmaf(c[1], kf[ky * 4 + 1], fmaf(c[2], kf[ky * 4 + 3], up[1]));
        up[2] = fmaf(c[1], kf[ky * 4 + 0], fmaf(c[2], kf[ky * 4 + 2], up[2]));
        up[3] = fmaf(c[2], kf[ky * 4 + 1], fmaf(c[3], kf[ky * 4 + 3], up[3]));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] += up[j];
    }
    *reinterpret_cast<float4*>(rgb + (size_t)i4 * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

int launch_rgb_combine(const float* partial, int num_partials, const float* bias, const float* skip,
                       const float* up_kernel, float* rgb, int B, int H, int W, cudaStream_t st) {
  HF_REQUIRE(rgb && (partial || num_partials == 0), "rgb_combine: null pointer");
  HF_REQUIRE(!skip || up_kernel, "rgb_combine: skip given without an upsample kernel");
  int64_t total = (int64_t)B * 3 * H * W;
  if (W % 4 == 0 && ((((uintptr_t)partial | (uintptr_t)rgb) & 15) == 0) && total / 4 < (int64_t)2000000000) {
    int64_t total4 = total / 4;
    int grid = (int)std::min<int64_t>((total4 + 255) / 256, (int64_t)num_sms() * 16);
    rgb_combine_vec4_kernel<<<grid, 256, 0, st>>>(partial, num_partials, total, bias, skip, up_kernel, rgb, H, W,
                                                  total4);
    HF_LAUNCH_OK("rgb_combine");
    count_launch();
    return HF_OK;
  }
  int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 16);
  rgb_combine_kernel<<<grid, 256, 0, st>>>(partial, num_partials, total, bias, skip, up_kernel, rgb, H, W, total);
  HF_LAUNCH_OK("rgb_combine");
  count_launch();
  return HF_OK;
}

// =============================================================================================
// standalone ToRGB on an fp32 NCHW feature map (module-level API): one pass over x
// =============================================================================================
__global__ void __launch_bounds__(256) torgb_nchw_kernel(const float* __restrict__ x, const float* __restrict__ w1,
                                                         float w_scale, const float* __restrict__ s,
                                                         float* __restrict__ part, int C, int HW) {
  extern __shared__ float ws[];      // [3][C] = w1[j,i] * scale * s[b,i]
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 3 * C; i += 256) {
    int c = i % C;
    ws[i] = __ldg(w1 + i) * w_scale * __ldg(s + (size_t)b * C + c);
  }
  __syncthreads();
  const float* xb = x + (size_t)b * C * HW;
  for (int p = blockIdx.x * 256 + threadIdx.x; p < HW; p += gridDim.x * 256) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
      float v = __ldg(xb + (size_t)c * HW + p);
      a0 = fmaf(v, ws[c], a0);
      a1 = fmaf(v, ws[C + c], a1);
      a2 = fmaf(v, ws[2 * C + c], a2);
    }
    float* pb = part + (size_t)b * 3 * HW;
    pb[p] = a0; pb[HW + p] = a1; pb[2 * (size_t)HW + p] = a2;
  }
}

int launch_torgb_nchw(const float* x, const float* w1, float w_scale, const float* s, const float* bias,
                      const float* skip, const float* up_kernel, float* y, int B, int C, int H, int W,
                      cudaStream_t st) {
  HF_REQUIRE(x && w1 && s && y, "torgb: null pointer");
  HF_REQUIRE(3 * C * sizeof(float) <= 48 * 1024, "torgb: too many channels (%d)", C);
  const int HW = H * W;
  dim3 grid(std::min(cdiv(HW, 256), num_sms() * 4), B);
  // y doubles as the partial buffer; rgb_combine then adds bias + upsampled skip in place.
  torgb_nchw_kernel<<<grid, 256, 3 * C * sizeof(float), st>>>(x, w1, w_scale, s, y, C, HW);
  HF_LAUNCH_OK("torgb_nchw");
  count_launch();
  return launch_rgb_combine(y, 1, bias, skip, up_kernel, y, B, H, W, st);
}

// =============================================================================================
// weight packing (once per weight load)
// =============================================================================================
// plain: wpk[o][tap*Cin + c] = W[o,c,dy,dx] * scale (tap = dy*3+dx, correlation form of F.conv2d)
// up   : the stride-2 transposed conv (model.py:260) composed with the 4x4 blur (pad (1,1), model.py:263)
//        as four 3x3 correlations, one per output parity (py,px); N row = (o/32)*128 + (py*2+px)*32 + o%32
//        Kp[py,px][dy,dx] = sum_{a,b} W[a,b] * blur[2-a+u, 2-b+v],  u = py+2-2dy, v = px+2-2dx
template <int DT>
__global__ void __launch_bounds__(256) pack_conv_kernel(const float* __restrict__ w, const float* __restrict__ blur,
                                                        uint16_t* __restrict__ wpk, float* __restrict__ wsq,
                                                        int Cout, int Cin, int ksize, int up, float scale) {
  const int taps = ksize * ksize;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)Cout * Cin) return;
  const int c = (int)(idx % Cin), o = (int)(idx / Cin);
  float wv[9];
  float sq = 0.f;
  for (int t = 0; t < taps; ++t) {
    wv[t] = w[((size_t)o * Cin + c) * taps + t] * scale;
    sq = fmaf(wv[t], wv[t], sq);
  }
  if (wsq) wsq[idx] = sq;
  const size_t K = (size_t)taps * Cin;
  if (!up) {
    for (int t = 0; t < taps; ++t) wpk[(size_t)o * K + (size_t)t * Cin + c] = Half2T<DT>::one(wv[t]);
  } else {
    float bk[16];
    for (int i = 0; i < 16; ++i) bk[i] = blur[i];
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        const size_t row = (size_t)(o / 32) * 128 + (size_t)(py * 2 + px) * 32 + (o % 32);
        for (int dy = 0; dy < 3; ++dy)
          for (int dx = 0; dx < 3; ++dx) {
            const int u = py + 2 - 2 * dy, v = px + 2 - 2 * dx;
            float acc = 0.f;
            for (int a = 0; a < 3; ++a)
              for (int b = 0; b < 3; ++b) {
                int iy = 2 - a + u, ix = 2 - b + v;
                if (iy >= 0 && iy < 4 && ix >= 0 && ix < 4) acc = fmaf(wv[a * 3 + b], bk[iy * 4 + ix], acc);
              }
            wpk[row * K + (size_t)(dy * 3 + dx) * Cin + c] = Half2T<DT>::one(acc);
          }
      }
  }
}

int launch_pack_conv(const float* w, const float* blur, void* wpk, float* wsq, int Cout, int Cin, int ksize,
                     int up, int nc, int dtype, cudaStream_t st) {
  (void)nc;
  HF_REQUIRE(w && wpk, "pack_conv: null pointer");
  HF_REQUIRE(ksize == 3 || ksize == 1, "pack_conv: kernel size %d unsupported", ksize);
  HF_REQUIRE(!up || (ksize == 3 && blur && Cout % 32 == 0), "pack_conv: upsample needs 3x3, blur kernel, Cout%%32==0");
  const float scale = 1.0f / sqrtf((float)(Cin * ksize * ksize));
  int64_t total = (int64_t)Cout * Cin;
  int grid = cdiv(total, 256);
  if (dtype == HF_BF16)
    pack_conv_kernel<HF_BF16><<<grid, 256, 0, st>>>(w, blur, (uint16_t*)wpk, wsq, Cout, Cin, ksize, up, scale);
  else
    pack_conv_kernel<HF_F16><<<grid, 256, 0, st>>>(w, blur, (uint16_t*)wpk, wsq, Cout, Cin, ksize, up, scale);
  HF_LAUNCH_OK("pack_conv");
  count_launch();
  return HF_OK;
}

__global__ void scale_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n, float scale) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[i] * scale;
}
int launch_scale_copy(const float* src, float* dst, int64_t n, float scale, cudaStream_t st) {
  HF_REQUIRE(src && dst, "scale_copy: null pointer");
  if (n == 0) return HF_OK;
  int grid = (int)std::min<int64_t>((n + 255) / 256, 4096);
  scale_copy_kernel<<<grid, 256, 0, st>>>(src, dst, n, scale);
  HF_LAUNCH_OK("scale_copy");
  count_launch();
  return HF_OK;
}

}  // namespace hf
