// tcgen05 implicit-GEMM convolution for sm_100a: the ModulatedConv2d / StyledConv hot loop.
//
//   y[b,o,p] = act( d[b,o] * sum_{tap,c} W~[o,tap,c] * xh[b, p+tap, c]  + nw*noise[b,p] + bias[o] )
//
// (reference: models/stylegan2/model.py:238-279 + :288-293 + op/fused_act.py:73-82, restated in the
// shared-weight form of SURVEY Appendix C-1: the per-sample style scale s[b,c] is folded into the
// 16-bit NHWC activation `xh` by the producer of that tensor, the demodulation d[b,o] is applied to
// the fp32 accumulator in the epilogue).  The upsampling conv (conv_transpose2d stride 2 + 4x4 blur,
// model.py:252-263) runs as four 3x3 correlations, one per output parity, stacked along GEMM-N
// (Appendix C-2), so the same kernel serves both.
//
// GEMM view per CTA tile:  D[128 pixels, n_tile] += A[128 pixels, 64 ch] * B[n_tile, 64 ch]^T over
// taps x channel chunks.  A = one TMA 4-D box (64ch, TW, TH, TB) of the NHWC activation shifted by
// the tap offset (out-of-bounds rows/cols are zero-filled by TMA = the conv padding); B = one TMA
// 2-D box of the packed weights.  Both land in 128B-swizzled K-major shared memory and feed
// tcgen05.mma (kind::f16, fp32 accumulators in TMEM).  Warp roles: warp 0 = TMA producer, warp 1 =
// MMA issuer (+ TMEM alloc), warps 2-5 = epilogue (TMEM -> registers -> demod/noise/bias/lrelu ->
// {16-bit NHWC for the next conv, fp32 NCHW, fused ToRGB partial sums}).  Persistent CTAs, static
// tile schedule, double-buffered accumulators so the epilogue of tile i overlaps the MMAs of i+1.
#include "hf_kernels.cuh"

namespace hf {

constexpr int kConvThreads = 192;
constexpr int kMaxStages = 8;
constexpr int kTableBytes = 16384;   // 512 entries x 32 B
constexpr float kSqrt2 = 1.41421356237309515f;

struct ConvKernelParams {
  int B, H, W, Cin, Cout;
  int Ho, Wo;
  int taps, up, act;
  int TW, TH, TB, tiles_x, tiles_y;
  int n_tile, num_n_tiles, num_tiles;
  int num_kb;              // taps * Cin / KCHUNK
  int kc_per_tap;          // Cin / KCHUNK
  int stages;
  uint32_t stage_bytes, a_bytes;
  uint32_t idesc;
  const float* d;
  const float* noise;
  int64_t noise_bstride;
  const float* noise_w;
  const float* bias;
  const float* s_next;
  uint16_t* xhat_out;
  float* out_nchw;
  const float* rgb_w;
  const float* rgb_s;
  float* rgb_partial;
};

struct __align__(16) TableEntry {
  float d, bias, s_next, pad;
  float w0, w1, w2, pad2;
};

template <int KCHUNK, int DT>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const ConvKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  TableEntry* table = reinterpret_cast<TableEntry*>(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(table) + kTableBytes);
  uint64_t* full_bar = bars;                       // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;         // [kMaxStages]
  uint64_t* tmem_full = bars + 2 * kMaxStages;     // [2]
  uint64_t* tmem_empty = bars + 2 * kMaxStages + 2;  // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t ROW_BYTES = KCHUNK * 2;
  constexpr uint32_t SBO = 8 * ROW_BYTES;
  constexpr uint32_t LAYOUT = (KCHUNK == 64) ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int pad = (p.taps == 9) ? 1 : 0;
      const uint32_t tx_bytes = p.stage_bytes;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int nt = tile % p.num_n_tiles, mt = tile / p.num_n_tiles;
        const int xt = mt % p.tiles_x, yt = (mt / p.tiles_x) % p.tiles_y, bt = mt / (p.tiles_x * p.tiles_y);
        const int x0 = xt * p.TW, y0 = yt * p.TH, b0 = bt * p.TB, n0 = nt * p.n_tile;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dy = (p.taps == 9) ? tap / 3 : 0, dx = (p.taps == 9) ? tap % 3 : 0;
          for (int kc = 0; kc < p.kc_per_tap; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* a_dst = stage_base + (size_t)stage * p.stage_bytes;
            uint8_t* b_dst = a_dst + p.a_bytes;
            mbar_expect_tx(&full_bar[stage], tx_bytes);
            tma_load_4d(a_dst, &tmA, &full_bar[stage], kc * KCHUNK, x0 + dx - pad, y0 + dy - pad, b0);
            tma_load_2d(b_dst, &tmB, &full_bar[stage], tap * p.Cin + kc * KCHUNK, n0);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(buf * 256);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(stage_base + (size_t)stage * p.stage_bytes);
          const uint32_t b_addr = a_addr + p.a_bytes;
#pragma unroll
          for (int k = 0; k < KCHUNK / 16; ++k) {
            const uint64_t adesc = make_kmajor_desc(a_addr + k * 32, SBO, LAYOUT);
            const uint64_t bdesc = make_kmajor_desc(b_addr + k * 32, SBO, LAYOUT);
            umma_f16(tmem_d, adesc, bdesc, p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                    // frees the smem slot when the MMAs retire
          if (kb == p.num_kb - 1) umma_commit(&tmem_full[buf]);   // accumulator ready for the epilogue
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ================================ epilogue (4 warps) ==========================
    const int wq = warp & 3;                 // TMEM lane quarter this warp may read
    const int row = wq * 32 + lane;          // GEMM row = pixel within the tile
    const int etid = (warp - 2) * 32 + lane; // 0..127
    const int w_l = row % p.TW, h_l = (row / p.TW) % p.TH, bb = row / (p.TW * p.TH);
    const float nw = p.noise_w ? __ldg(p.noise_w) : 0.f;
    const int chunks = p.n_tile / 32;
    const int nc_tile = p.up ? p.n_tile / 4 : p.n_tile;   // output channels per tile
    const size_t plane_o = (size_t)p.Ho * p.Wo;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int nt = tile % p.num_n_tiles, mt = tile / p.num_n_tiles;
      const int xt = mt % p.tiles_x, yt = (mt / p.tiles_x) % p.tiles_y, bt = mt / (p.tiles_x * p.tiles_y);
      const int x = xt * p.TW + w_l, y = yt * p.TH + h_l, b = bt * p.TB + bb;
      const int n0 = nt * p.n_tile;
      const bool valid = b < p.B;
      // ---- per-tile table: demod, bias, next-layer style scale, ToRGB weights, indexed [bb][col]
      // (one entry per output channel of the tile: the four parity column groups of an up-conv share it)
      for (int e = etid; e < p.TB * nc_tile; e += 128) {
        const int ebb = e / nc_tile, ol = e - ebb * nc_tile;
        const int eb = bt * p.TB + ebb;
        const int o = (p.up ? (n0 >> 2) : n0) + ol;
        TableEntry t;
        t.d = 1.f; t.bias = 0.f; t.s_next = 1.f; t.pad = 0.f; t.w0 = t.w1 = t.w2 = 0.f; t.pad2 = 0.f;
        if (eb < p.B) {
          const size_t bo = (size_t)eb * p.Cout + o;
          if (p.d) t.d = __ldg(p.d + bo);
          if (p.bias) t.bias = __ldg(p.bias + o);
          if (p.s_next) t.s_next = __ldg(p.s_next + bo);
          if (p.rgb_w) {
            const float rs = __ldg(p.rgb_s + bo);
            t.w0 = __ldg(p.rgb_w + o) * rs;
            t.w1 = __ldg(p.rgb_w + p.Cout + o) * rs;
            t.w2 = __ldg(p.rgb_w + 2 * p.Cout + o) * rs;
          }
        }
        table[e] = t;
      }
      // ---- noise for this pixel (plain) / its 2x2 output quad (up)
      float nz[4] = {0.f, 0.f, 0.f, 0.f};
      if (p.noise && valid) {
        const float* np_ = p.noise + (size_t)b * p.noise_bstride;
        if (p.up) {
#pragma unroll
          for (int par = 0; par < 4; ++par)
            nz[par] = nw * __ldg(np_ + (size_t)(2 * y + (par >> 1)) * p.Wo + 2 * x + (par & 1));
        } else {
          nz[0] = nw * __ldg(np_ + (size_t)y * p.Wo + x);
        }
      }
      named_bar_sync(1, 128);

      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait(&tmem_full[buf], use & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(buf * 256);
      const TableEntry* trow = table + bb * nc_tile;
      float rgb[4][3];
#pragma unroll
      for (int i = 0; i < 4; ++i) rgb[i][0] = rgb[i][1] = rgb[i][2] = 0.f;

      for (int q = 0; q < chunks; ++q) {
        uint32_t acc[32];
        tmem_ld_32x32(taddr + q * 32, acc);
        tmem_ld_wait();
        if (q == chunks - 1) {     // accumulator fully read: hand the TMEM buffer back to the MMA warp
          tc_fence_before();
          mbar_arrive(&tmem_empty[buf]);
        }
        const int par = p.up ? (q & 3) : 0;
        const int t_base = p.up ? (q >> 2) * 32 : q * 32;          // channel offset inside the tile
        const int o_base = (p.up ? (n0 >> 2) : n0) + t_base;
        const int yo = p.up ? 2 * y + (par >> 1) : y;
        const int xo = p.up ? 2 * x + (par & 1) : x;
        const float nzv = nz[par];
        uint32_t packed[16];
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
        float* onchw = (p.out_nchw && valid) ? p.out_nchw + ((size_t)b * p.Cout + o_base) * plane_o +
                                                  (size_t)yo * p.Wo + xo
                                                : nullptr;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float v[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const TableEntry& t = trow[t_base + j + u];
            float a = fmaf(__uint_as_float(acc[j + u]), t.d, nzv + t.bias);
            if (p.act) a = (a > 0.f ? a : 0.2f * a) * kSqrt2;
            r0 = fmaf(a, t.w0, r0); r1 = fmaf(a, t.w1, r1); r2 = fmaf(a, t.w2, r2);
            if (onchw) onchw[(size_t)(j + u) * plane_o] = a;
            v[u] = a * t.s_next;
          }
          packed[j >> 1] = Half2T<DT>::pack(v[0], v[1]);
        }
        rgb[par][0] += r0; rgb[par][1] += r1; rgb[par][2] += r2;
        if (p.xhat_out && valid) {
          uint4* dst = reinterpret_cast<uint4*>(p.xhat_out + (((size_t)b * p.Ho + yo) * p.Wo + xo) * p.Cout + o_base);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
        }
      }
      if (p.rgb_partial && valid) {
        float* pp = p.rgb_partial + ((size_t)nt * p.B + b) * 3 * plane_o;
        const int npar = p.up ? 4 : 1;
        for (int par = 0; par < npar; ++par) {
          const int yo = p.up ? 2 * y + (par >> 1) : y;
          const int xo = p.up ? 2 * x + (par & 1) : x;
#pragma unroll
          for (int j = 0; j < 3; ++j) pp[(size_t)j * plane_o + (size_t)yo * p.Wo + xo] = rgb[par][j];
        }
      }
      named_bar_sync(1, 128);    // table is rewritten by the next tile
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int conv_up_nc(int Cout) { return Cout < 32 ? Cout : 32; }

int conv_plan(const ConvLaunch& a, ConvPlan* p) {
  HF_REQUIRE(a.B > 0 && a.H > 0 && a.W > 0, "conv: bad shape B=%d H=%d W=%d", a.B, a.H, a.W);
  HF_REQUIRE(a.taps == 9 || a.taps == 1, "conv: taps must be 9 or 1");
  HF_REQUIRE(a.Cin % 32 == 0 && a.Cin >= 32, "conv: Cin=%d must be a multiple of 32", a.Cin);
  HF_REQUIRE(a.Cout % 32 == 0 && a.Cout >= 32, "conv: Cout=%d must be a multiple of 32", a.Cout);
  HF_REQUIRE(!a.up || a.taps == 9, "conv: upsample needs a 3x3 kernel");
  p->TW = a.W < 16 ? a.W : 16;
  HF_REQUIRE(128 % p->TW == 0, "conv: width %d unsupported (tile width must divide 128)", a.W);
  p->TH = a.H < 128 / p->TW ? a.H : 128 / p->TW;
  HF_REQUIRE(128 % (p->TW * p->TH) == 0, "conv: %dx%d image does not tile into 128 GEMM rows", a.H, a.W);
  p->TB = 128 / (p->TW * p->TH);
  HF_REQUIRE(a.W % p->TW == 0 && a.H % p->TH == 0, "conv: %dx%d not divisible by tile %dx%d", a.H, a.W, p->TH, p->TW);
  p->tiles_x = a.W / p->TW;
  p->tiles_y = a.H / p->TH;
  p->tiles_b = (a.B + p->TB - 1) / p->TB;
  p->num_m_tiles = p->tiles_x * p->tiles_y * p->tiles_b;
  p->kchunk = (a.Cin % 64 == 0) ? 64 : 32;
  const int ntot = a.up ? 4 * a.Cout : a.Cout;
  const int nmin = a.up ? 128 : 32;
  const int sms = num_sms();
  int n_tile = 0;
  if (a.force_n_tile) {
    n_tile = a.force_n_tile;
  } else {
    for (int cand = 256; cand >= nmin; cand >>= 1) {
      if (ntot % cand || p->TB * (a.up ? cand / 4 : cand) > 512) continue;
      n_tile = cand;
      if ((int64_t)p->num_m_tiles * (ntot / cand) >= sms) break;   // largest tile that still fills the GPU
    }
  }
  HF_REQUIRE(n_tile >= nmin && n_tile <= 256 && ntot % n_tile == 0 && n_tile % 32 == 0 &&
                 p->TB * (a.up ? n_tile / 4 : n_tile) <= 512,
             "conv: no valid N tile (Ntot=%d, n_tile=%d, TB=%d)", ntot, n_tile, p->TB);
  p->n_tile = n_tile;
  p->num_n_tiles = ntot / n_tile;
  p->nc = a.up ? n_tile / 4 : n_tile;
  const size_t stage_bytes = (size_t)128 * p->kchunk * 2 + (size_t)n_tile * p->kchunk * 2;
  const size_t fixed = 1024 /*align*/ + kTableBytes + 512 /*barriers*/;
  int stages = (int)((232448 - fixed) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  HF_REQUIRE(stages >= 2, "conv: not enough shared memory for 2 stages");
  p->stages = stages;
  p->smem_bytes = fixed + stages * stage_bytes;
  p->num_tiles = p->num_m_tiles * p->num_n_tiles;
  p->grid = p->num_tiles < sms ? p->num_tiles : sms;
  return HF_OK;
}

template <int KCHUNK, int DT>
static int launch_conv_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvKernelParams& kp,
                         const ConvPlan& pl, cudaStream_t st) {
  static bool attr_set = false;
  auto kern = conv_igemm_kernel<KCHUNK, DT>;
  if (!attr_set) {
    HF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  kern<<<pl.grid, kConvThreads, pl.smem_bytes, st>>>(tmA, tmB, kp);
  HF_LAUNCH_OK("conv_igemm");
  count_launch();
  return HF_OK;
}

int launch_conv(const ConvLaunch& a, cudaStream_t st, ConvPlan* plan_out) {
  ConvPlan pl;
  int rc = conv_plan(a, &pl);
  if (rc) return rc;
  HF_REQUIRE(a.xhat_in && a.wpk, "conv: null operand pointer");
  HF_REQUIRE(!a.rgb_partial || (a.rgb_w && a.rgb_s), "conv: rgb_partial needs rgb_w and rgb_s");
  HF_REQUIRE(!a.noise || a.noise_w, "conv: noise given without noise weight");
  HF_REQUIRE((((uintptr_t)a.xhat_in | (uintptr_t)a.wpk | (uintptr_t)a.xhat_out) & 15) == 0,
             "conv: 16-bit tensors must be 16-byte aligned");

  alignas(64) CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)a.Cin, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B};
    uint64_t strides[3] = {(uint64_t)a.Cin * 2, (uint64_t)a.W * a.Cin * 2, (uint64_t)a.H * a.W * a.Cin * 2};
    uint32_t box[4] = {(uint32_t)pl.kchunk, (uint32_t)pl.TW, (uint32_t)pl.TH, (uint32_t)pl.TB};
    rc = encode_tmap(&tmA, a.dtype, 4, const_cast<void*>(a.xhat_in), dims, strides, box, pl.kchunk * 2);
    if (rc) return rc;
  }
  {
    const uint64_t K = (uint64_t)a.taps * a.Cin;
    const uint64_t N = a.up ? 4ull * a.Cout : (uint64_t)a.Cout;
    uint64_t dims[2] = {K, N};
    uint64_t strides[1] = {K * 2};
    uint32_t box[2] = {(uint32_t)pl.kchunk, (uint32_t)pl.n_tile};
    rc = encode_tmap(&tmB, a.dtype, 2, const_cast<void*>(a.wpk), dims, strides, box, pl.kchunk * 2);
    if (rc) return rc;
  }

  ConvKernelParams kp;
  kp.B = a.B; kp.H = a.H; kp.W = a.W; kp.Cin = a.Cin; kp.Cout = a.Cout;
  kp.Ho = a.up ? 2 * a.H : a.H;
  kp.Wo = a.up ? 2 * a.W : a.W;
  kp.taps = a.taps; kp.up = a.up; kp.act = a.act;
  kp.TW = pl.TW; kp.TH = pl.TH; kp.TB = pl.TB; kp.tiles_x = pl.tiles_x; kp.tiles_y = pl.tiles_y;
  kp.n_tile = pl.n_tile; kp.num_n_tiles = pl.num_n_tiles; kp.num_tiles = pl.num_tiles;
  kp.kc_per_tap = a.Cin / pl.kchunk;
  kp.num_kb = a.taps * kp.kc_per_tap;
  kp.stages = pl.stages;
  kp.a_bytes = 128u * pl.kchunk * 2;
  kp.stage_bytes = kp.a_bytes + (uint32_t)pl.n_tile * pl.kchunk * 2;
  kp.idesc = make_idesc_f16(a.dtype, 128, pl.n_tile);
  kp.d = a.d;
  kp.noise = a.noise;
  kp.noise_bstride = (a.noise && a.noise_batch > 1) ? (int64_t)kp.Ho * kp.Wo : 0;
  kp.noise_w = a.noise ? a.noise_w : nullptr;
  kp.bias = a.bias;
  kp.s_next = a.s_next;
  kp.xhat_out = (uint16_t*)a.xhat_out;
  kp.out_nchw = a.out_nchw;
  kp.rgb_w = a.rgb_partial ? a.rgb_w : nullptr;
  kp.rgb_s = a.rgb_s;
  kp.rgb_partial = a.rgb_partial;
  if (plan_out) *plan_out = pl;

  if (pl.kchunk == 64) {
    return a.dtype == HF_BF16 ? launch_conv_t<64, HF_BF16>(tmA, tmB, kp, pl, st)
                              : launch_conv_t<64, HF_F16>(tmA, tmB, kp, pl, st);
  }
  return a.dtype == HF_BF16 ? launch_conv_t<32, HF_BF16>(tmA, tmB, kp, pl, st)
                            : launch_conv_t<32, HF_F16>(tmA, tmB, kp, pl, st);
}

}  // namespace hf
