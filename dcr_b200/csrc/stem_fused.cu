// ResNet stem (7x7 / stride 2 / pad 3 convolution + BatchNorm + ReLU) as a tcgen05 implicit GEMM whose A operand is an
// OVERLAPPING-WINDOW (Toeplitz) view of the input in shared memory -- every input pixel travels L2 -> shared memory ~1.7
// times instead of 16 times (conv_gemm.cu's space-to-depth path re-reads each stored pixel once per window position and
// filter row: 357 us of a 4.07 ms SSCD forward at batch 256, ingest bound; profiles/r01_layers_sscd.txt).
//
// Reference call site: `model(samples)` (utils_ret.py:751) -> torchvision ResNet conv1 / bn1 / relu of the SSCD trunk.
//
// Input layout (stem_rows_u8_kernel, fused with Resize/CenterCrop/ToTensor/Normalize of diff_retrieval.py:325-330):
// with ip the zero-padded (3 pixels) normalised crop, the image is stored as two "column-parity planes" of 16-byte units
//     plane_e[P * PW + u] = { ip[2P + i][2u + e][c] : i in {0,1}, c in {0,1,2} } + 2 zero channels      (8 bf16)
// so that   out[y][x] = sum_{a<4, e<2, b<4, ch<8} W[a][e][b][ch] * plane_e[(y + a) * PW + (x + b)][ch]
// (filter row 2a+i, filter column 2b+e; the 8th row / column of the 8x8 footprint carries zero weights).  With output
// position m = y * PW + x the A operand of K-chunk (a, e, b) is the SAME linear array shifted by (a * PW + b) units: in a
// K-major SWIZZLE_NONE shared-memory descriptor rows are 16 bytes apart (stride-dimension offset 128 B per 8 rows) and
// the second 16-byte K chunk of an instruction sits leading-dimension-offset = 16 bytes further -- i.e. row m+1 and
// K-chunk b+1 address the same bytes.  tools/microbench/toeplitz_probe.cu verifies the hardware accepts this.
// Positions with x >= OW (PW - OW per row) are junk and dropped by the epilogue.
//
// Roles (352 threads, persistent over tiles of 512 positions = 4 MMA row blocks):
//   warp 0   producer: two cp.async.bulk copies per tile (the even / odd plane windows, 512 + 3*PW + 3 units each)
//   warps 1, 10   tcgen05.mma issuers (alternate row blocks): 16 x (128 x 64 x 16) per row block, weights (64 x 256, 32 KB,
//            128B swizzle) resident
//   warps 2-9 epilogue: TMEM -> BN affine + ReLU -> bf16 -> staging -> coalesced NHWC stores of the valid positions
#include <cuda_bf16.h>

#include <algorithm>
#include <cstring>

#include "dcr_internal.cuh"
#include "host_util.cuh"
#include "ptx.cuh"

namespace dcr {

namespace {

constexpr int kSN = 64;                   // output channels
constexpr int kSTile = 512;               // positions per tile
constexpr int kSBlocks = kSTile / 128;    // MMA row blocks per tile
constexpr int kSThreads = 352;             // warp 0 producer, warps 1 and 10 MMA issuers, warps 2-9 epilogue
constexpr int kWBytes = kSN * 256 * 2;    // resident weights: 4 k-blocks of [64 rows x 128 B]

struct StemParams {
  const __nv_bfloat16* planes;   // [B][2][alloc_units][8]
  long long img_stride;          // elements between images (2 * alloc_units * 8)
  long long plane_stride;        // elements between the two planes (alloc_units * 8)
  int B, OH, OW, PW;
  // work units: a unit is `tiles_per_unit` consecutive 512-position tiles of one image, starting at position
  // unit_begin(part).  Without pooling: unit = one tile.  With pooling: unit = 1/parts of an image (one extra conv row on top).
  int units_per_img, tiles_per_unit, num_units;
  int part_rows;                 // pooling: conv rows owned per part (2 * pooled rows per part)
  int OHp, OWp, prow_per_part;   // pooling: pooled output size, pooled rows per part
  int win_units;                 // units copied per plane and tile: 512 + 3 * PW + 3, rounded up to 8
  int win_stages;
  const float* scale;
  const float* bias;
  __nv_bfloat16* out;            // NHWC [B][OH][OW][64]
};

DCR_DEVICE uint64_t desc_nosw(uint32_t addr) {
  // K-major, SWIZZLE_NONE: 8-row x 16-byte core matrices; LBO (next K chunk) = 16 B, SBO (next 8 rows) = 128 B
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(16 >> 4) << 16;
  d |= static_cast<uint64_t>(128 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

DCR_DEVICE void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

DCR_DEVICE uint32_t pack2s(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

// first position of a unit
DCR_DEVICE int unit_begin(const StemParams& p, int part, bool pool) {
  if (!pool) return part * kSTile;
  const int row_lo = max(0, part * p.part_rows - 1);     // one conv row above the part's first pooled window
  return row_lo * p.PW;
}

// kPool: the 3x3 / stride 2 / pad 1 max pool that follows the stem (torchvision ResNet.maxpool) is taken in the epilogue:
// conv outputs (post BN + ReLU, bf16) go into a ring of 8 conv rows in shared memory and a pooled row is emitted as soon
// as its three conv rows are complete -- the 112 x 112 x 64 stem activation (411 MB at batch 256) never reaches HBM.
template <bool kPool>
__global__ void __launch_bounds__(kSThreads, 1) stem_conv_kernel(const __grid_constant__ CUtensorMap tmap_w, const StemParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int win_bytes = p.win_units * 16;                       // per plane
  const int stage_bytes = (2 * win_bytes + 1023) & ~1023;
  uint8_t* s_w = smem;                                          // 32 KB, 4 k-blocks
  uint8_t* s_win = s_w + kWBytes;                               // win_stages x [even | odd]
  uint8_t* s_out = s_win + p.win_stages * stage_bytes;          // 2 x [128 positions x 128 B] | kPool: ring of 8 conv rows
  const int ring_row_bytes = p.OW * 128;
  float* sb = reinterpret_cast<float*>(s_out + (kPool ? ((8 * ring_row_bytes + 1023) & ~1023) : 2 * 16384));   // scale[64] | bias[64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sb + 128);
  uint64_t* w_full = bars;
  uint64_t* win_full = bars + 1;      // [4]
  uint64_t* win_empty = bars + 5;     // [4]
  uint64_t* t_full = bars + 9;        // [2][4]
  uint64_t* t_empty = bars + 17;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_w);
  if (warp == 1 && lane == 0) {
    mbar_init(w_full, 1);
    for (int s = 0; s < 4; ++s) {
      mbar_init(&win_full[s], 1);
      mbar_init(&win_empty[s], 2);   // both MMA issuers commit once per tile
    }
    for (int s = 0; s < 8; ++s) mbar_init(&t_full[s], 1);
    for (int s = 0; s < 2; ++s) mbar_init(&t_empty[s], 8);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<1>(tmem_slot, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== producer =====================================
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, kWBytes);
      for (int kb = 0; kb < 4; ++kb) tma_load_2d<1>(s_w + kb * 8192, &tmap_w, w_full, kb * 64, 0, kEvictLast);
    }
    __syncwarp();
    PipeState ws(p.win_stages);
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      const int b = unit / p.units_per_img;
      const int mu = unit_begin(p, unit - b * p.units_per_img, kPool);
      for (int t = 0; t < p.tiles_per_unit; ++t, ws.next()) {
        const int m0 = mu + t * kSTile;
        mbar_wait(&win_empty[ws.s], ws.ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&win_full[ws.s], 2 * win_bytes);
          const __nv_bfloat16* src = p.planes + static_cast<size_t>(b) * p.img_stride + static_cast<size_t>(m0) * 8;
          bulk_load(s_win + ws.s * stage_bytes, src, win_bytes, &win_full[ws.s]);
          bulk_load(s_win + ws.s * stage_bytes + win_bytes, src + p.plane_stride, win_bytes, &win_full[ws.s]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ===================================== MMA issuers =====================================
    // Two issuing warps on separate accumulators: one thread issues a 128x64x16 tcgen05.mma every ~90 cycles at best (the
    // tensor core needs 32), so the row blocks of a tile alternate between two issuers (tools/microbench/umma_rate.cu).
    const int issuer = (warp == 1) ? 0 : 1;
    constexpr uint32_t idesc = umma_idesc_bf16(128, kSN);
    mbar_wait(w_full, 0);
    tc_fence_after();
    const uint64_t dw0 = umma_desc_sw128(smem_u32(s_w));
    const uint32_t win0 = smem_u32(s_win);
    PipeState ws(p.win_stages);
    uint32_t tc = 0;
    const uint32_t PW = static_cast<uint32_t>(p.PW);
    int my_tiles = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) my_tiles += p.tiles_per_unit;
    for (int it = 0; it < my_tiles; ++it, ++tc, ws.next()) {
      const uint32_t buf = tc & 1;
      mbar_wait(&win_full[ws.s], ws.ph);
      mbar_wait(&t_empty[buf], ((tc >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t wbase = win0 + ws.s * stage_bytes;
#pragma unroll 1
      for (int mb = issuer; mb < kSBlocks; mb += 2) {
        const uint32_t tmem_d = tmem_base + (buf * kSBlocks + mb) * kSN;
        if (elect_one()) {
#pragma unroll
          for (int a = 0; a < 4; ++a) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
#pragma unroll
              for (int bp = 0; bp < 2; ++bp) {
                // A: positions mb*128.., shifted by filter row pair a and column pair 2*bp (units of 16 B)
                const uint32_t a_addr = wbase + e * win_bytes + (mb * 128 + a * PW + 2 * bp) * 16;
                const int kc = (a * 2 + e) * 2 + bp;                 // K = 16 chunk of the weights
                const uint64_t db = dw0 + static_cast<uint64_t>((kc >> 2) * (8192 >> 4) + (kc & 3) * 2);
                umma_f16<1>(tmem_d, desc_nosw(a_addr), db, idesc, kc != 0);
              }
            }
          }
          umma_commit<1>(&t_full[buf * kSBlocks + mb]);
          if (mb + 2 >= kSBlocks) umma_commit<1>(&win_empty[ws.s]);   // this issuer's last row block of the tile
        }
        __syncwarp();
      }
    }
  } else if (warp < 10) {
    // ===================================== epilogue warps =====================================
    const uint32_t ewarp = warp - 2;
    const uint32_t quad = warp & 3;
    const uint32_t half = ewarp >> 2;                 // 32-column half of the 64 channels
    const uint32_t row = quad * 32 + lane;            // position inside the row block
    const uint32_t etid = ewarp * 32 + lane;
    const uint32_t tmem_row = tmem_base + ((quad * 32u) << 16);
    const uint32_t sb_addr = smem_u32(sb), so_addr = smem_u32(s_out);
    for (int c = etid; c < kSN; c += 256) {
      st_shared_f32(sb_addr + c * 4, p.scale ? p.scale[c] : 1.f);
      st_shared_f32(sb_addr + (kSN + c) * 4, p.bias ? p.bias[c] : 0.f);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    float sc[32], bi[32];
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
      const float4 s4 = ld_shared_f4(sb_addr + (half * 32 + c) * 4);
      const float4 b4 = ld_shared_f4(sb_addr + (kSN + half * 32 + c) * 4);
      sc[c] = s4.x; sc[c + 1] = s4.y; sc[c + 2] = s4.z; sc[c + 3] = s4.w;
      bi[c] = b4.x; bi[c + 1] = b4.y; bi[c + 2] = b4.z; bi[c + 3] = b4.w;
    }
    const int positions = p.OH * p.PW;
    uint32_t tc = 0, blk = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      const int b = unit / p.units_per_img;
      const int part = unit - b * p.units_per_img;
      const int mu = unit_begin(p, part, kPool);
      // pooling state of this unit: conv rows [row_lo, row_hi) are produced here, pooled rows [next_yp, yp_end) emitted
      const int row_lo = kPool ? max(0, part * p.part_rows - 1) : 0;
      const int row_hi = kPool ? min(p.OH, (part + 1) * p.part_rows) : p.OH;
      int next_yp = part * p.prow_per_part;
      const int yp_end = min(p.OHp, next_yp + p.prow_per_part);
      if constexpr (kPool) asm volatile("bar.sync 2, 256;" ::: "memory");   // nobody still pools the previous unit's rows
      for (int t = 0; t < p.tiles_per_unit; ++t, ++tc) {
        const int m0 = mu + t * kSTile;
        const uint32_t buf = tc & 1;
#pragma unroll 1
        for (int mb = 0; mb < kSBlocks; ++mb, ++blk) {
          mbar_wait(&t_full[buf * kSBlocks + mb], (tc >> 1) & 1);
          tc_fence_after();
          uint32_t r[32];
          tmem_ld_32x32(tmem_row + (buf * kSBlocks + mb) * kSN + half * 32, r);
          tmem_ld_wait_regs(r);
          if (mb == kSBlocks - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[buf]);
          }
          uint4 v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            v[q].x = pack2s(fmaxf(fmaf(__uint_as_float(r[q * 8 + 0]), sc[q * 8 + 0], bi[q * 8 + 0]), 0.f),
                            fmaxf(fmaf(__uint_as_float(r[q * 8 + 1]), sc[q * 8 + 1], bi[q * 8 + 1]), 0.f));
            v[q].y = pack2s(fmaxf(fmaf(__uint_as_float(r[q * 8 + 2]), sc[q * 8 + 2], bi[q * 8 + 2]), 0.f),
                            fmaxf(fmaf(__uint_as_float(r[q * 8 + 3]), sc[q * 8 + 3], bi[q * 8 + 3]), 0.f));
            v[q].z = pack2s(fmaxf(fmaf(__uint_as_float(r[q * 8 + 4]), sc[q * 8 + 4], bi[q * 8 + 4]), 0.f),
                            fmaxf(fmaf(__uint_as_float(r[q * 8 + 5]), sc[q * 8 + 5], bi[q * 8 + 5]), 0.f));
            v[q].w = pack2s(fmaxf(fmaf(__uint_as_float(r[q * 8 + 6]), sc[q * 8 + 6], bi[q * 8 + 6]), 0.f),
                            fmaxf(fmaf(__uint_as_float(r[q * 8 + 7]), sc[q * 8 + 7], bi[q * 8 + 7]), 0.f));
          }
          const int mblk = m0 + mb * 128;
          if constexpr (!kPool) {
            // staging row = position inside the block, 128 B per position, 16-byte chunks XOR-swizzled by the row
            const uint32_t stage = so_addr + (blk & 1) * 16384;
            const uint32_t srow = stage + row * 128;
            const uint32_t sw = row & 7;
#pragma unroll
            for (int q = 0; q < 4; ++q) st_shared_v4(srow + (((half * 4 + q) ^ sw) << 4), v[q]);
            asm volatile("bar.sync 2, 256;" ::: "memory");
            // coalesced copy-out: 8 threads per position (16 B each), 32 positions per pass; junk positions are skipped
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int pos = it * 32 + static_cast<int>(etid >> 3);
              const int ch16 = static_cast<int>(etid & 7);
              const int m = mblk + pos;
              const int y = m / p.PW, x = m - y * p.PW;
              if (m < positions && x < p.OW) {
                const uint4 o = ld_shared_v4(stage + pos * 128 + ((ch16 ^ (pos & 7)) << 4));
                __nv_bfloat16* dst = p.out + ((static_cast<size_t>(b) * p.OH + y) * p.OW + x) * kSN + ch16 * 8;
                *reinterpret_cast<uint4*>(dst) = o;
              }
            }
            // the staging buffer (blk & 1) is rewritten two blocks later: the barrier of the next block orders that
          } else {
            // conv row ring: ring[(y & 7)][x][64 ch], 16-byte chunks XOR-swizzled by x
            const int m = mblk + static_cast<int>(row);
            const int y = m / p.PW, x = m - y * p.PW;
            if (x < p.OW && y >= row_lo && y < row_hi) {
              const uint32_t rrow = so_addr + (y & 7) * ring_row_bytes + x * 128;
#pragma unroll
              for (int q = 0; q < 4; ++q) st_shared_v4(rrow + (((half * 4 + q) ^ (x & 7)) << 4), v[q]);
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");
            // rows <= yc are complete; emit every pooled row whose last conv row is in (ring depth 8 keeps the rows a slower
            // thread is still pooling apart from the rows the next block writes)
            const int yc = min((mblk + 128) / p.PW - 1, row_hi - 1);
            while (next_yp < yp_end && min(2 * next_yp + 1, p.OH - 1) <= yc) {
              // taps outside the image are replaced by the nearest tap INSIDE the window (the maximum is idempotent), so
              // there is no branching and the nine 16-byte loads of an item are independent and all in flight together
              const int y1 = 2 * next_yp;
              const uint32_t r0 = so_addr + (max(y1 - 1, 0) & 7) * ring_row_bytes;
              const uint32_t r1 = so_addr + (y1 & 7) * ring_row_bytes;
              const uint32_t r2 = so_addr + (min(y1 + 1, p.OH - 1) & 7) * ring_row_bytes;
              for (int idx = etid; idx < p.OWp * 8; idx += 256) {
                const int xp = idx >> 3, c = idx & 7;
                const int x1 = 2 * xp, x0 = max(x1 - 1, 0), x2 = min(x1 + 1, p.OW - 1);
                const uint32_t o0 = x0 * 128 + ((c ^ (x0 & 7)) << 4), o1 = x1 * 128 + ((c ^ (x1 & 7)) << 4),
                               o2 = x2 * 128 + ((c ^ (x2 & 7)) << 4);
                uint4 w[9];
                w[0] = ld_shared_v4(r0 + o0); w[1] = ld_shared_v4(r0 + o1); w[2] = ld_shared_v4(r0 + o2);
                w[3] = ld_shared_v4(r1 + o0); w[4] = ld_shared_v4(r1 + o1); w[5] = ld_shared_v4(r1 + o2);
                w[6] = ld_shared_v4(r2 + o0); w[7] = ld_shared_v4(r2 + o1); w[8] = ld_shared_v4(r2 + o2);
                uint4 o;
                {
                  auto mx = [](uint32_t a, uint32_t b) {
                    __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
                    return *reinterpret_cast<uint32_t*>(&r);
                  };
                  auto mx9 = [&](auto get) {
                    const uint32_t a = mx(mx(get(w[0]), get(w[1])), mx(get(w[2]), get(w[3])));
                    const uint32_t b = mx(mx(get(w[4]), get(w[5])), mx(get(w[6]), get(w[7])));
                    return mx(mx(a, b), get(w[8]));
                  };
                  o.x = mx9([](const uint4& v) { return v.x; });
                  o.y = mx9([](const uint4& v) { return v.y; });
                  o.z = mx9([](const uint4& v) { return v.z; });
                  o.w = mx9([](const uint4& v) { return v.w; });
                }
                __nv_bfloat16* dst = p.out + ((static_cast<size_t>(b) * p.OHp + next_yp) * p.OWp + xp) * kSN + c * 8;
                *reinterpret_cast<uint4*>(dst) = o;
              }
              ++next_yp;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

// ---- input kernel: uint8 HWC (or fp32 NCHW) image -> the two column-parity planes ----------------------------------------
struct StemRowsParams {
  const uint8_t* img;
  const float* img_f32;
  int B, IH, IW, crop_y, crop_x, H, W, RH, RW;
  float rscale;
  float mean[3], std[3], post_scale, post_shift;
  __nv_bfloat16* out;
  long long img_stride, plane_stride;
  int PW, rows;     // units per pair-row, pair-rows written (OH + 3)
};

template <bool kResize, bool kF32>
__global__ void __launch_bounds__(256) stem_rows_kernel(const StemRowsParams p) {
  __shared__ float lut[3][256];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    const int c = i >> 8, u = i & 255;
    const float val = (static_cast<float>(u) / 255.f - p.mean[c]) / p.std[c];   // ToTensor + Normalize, IEEE fp32
    lut[c][u] = p.post_scale * val + p.post_shift;
  }
  __syncthreads();
  const long long per_img = 2ll * p.rows * p.PW;
  const long long total = per_img * p.B;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(idx / per_img);
    long long rem = idx - b * per_img;
    const int e = static_cast<int>(rem / (static_cast<long long>(p.rows) * p.PW));
    rem -= static_cast<long long>(e) * p.rows * p.PW;
    const int P = static_cast<int>(rem / p.PW), u = static_cast<int>(rem % p.PW);
    const uint8_t* img = p.img + static_cast<size_t>(b) * p.IH * p.IW * 3;
    const size_t plane = static_cast<size_t>(p.IH) * p.IW;
    const float* imgf = p.img_f32 + static_cast<size_t>(b) * 3 * plane;
    auto px = [&](int yy, int xx, int c) -> float {
      if constexpr (kF32) return fmaf(p.post_scale, imgf[c * plane + static_cast<size_t>(yy + p.crop_y) * p.IW + (xx + p.crop_x)], p.post_shift);
      else return lut[c][img[(static_cast<size_t>(yy + p.crop_y) * p.IW + (xx + p.crop_x)) * 3 + c]];
    };
    float z[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) z[k] = 0.f;
    const int x = 2 * u + e - 3;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int y = 2 * P + i - 3;
      if (y < 0 || y >= p.RH || x < 0 || x >= p.RW) continue;
      if constexpr (!kResize) {
#pragma unroll
        for (int c = 0; c < 3; ++c) z[i * 3 + c] = px(y, x, c);
      } else {
        const float sy = fmaxf(p.rscale * (static_cast<float>(y) + 0.5f) - 0.5f, 0.f);
        const float sx = fmaxf(p.rscale * (static_cast<float>(x) + 0.5f) - 0.5f, 0.f);
        const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
        const int y1 = y0 + (y0 < p.H - 1 ? 1 : 0), x1 = x0 + (x0 < p.W - 1 ? 1 : 0);
        const float ly = sy - static_cast<float>(y0), lx = sx - static_cast<float>(x0);
        const float hy = 1.f - ly, hx = 1.f - lx;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          z[i * 3 + c] = hy * (hx * px(y0, x0, c) + lx * px(y0, x1, c)) + ly * (hx * px(y1, x0, c) + lx * px(y1, x1, c));
      }
    }
    uint4 v;
    v.x = pack2s(z[0], z[1]);
    v.y = pack2s(z[2], z[3]);
    v.z = pack2s(z[4], z[5]);
    v.w = pack2s(z[6], z[7]);
    __nv_bfloat16* dst = p.out + static_cast<size_t>(b) * p.img_stride + static_cast<size_t>(e) * p.plane_stride +
                         (static_cast<size_t>(P) * p.PW + u) * 8;
    *reinterpret_cast<uint4*>(dst) = v;
  }
}

}  // namespace

// geometry shared by the host graph builder (through dcr_stem_plane_units) and the two launchers
int stem_fused_pitch(int out_w) { return out_w + 4; }
long long stem_fused_plane_units(int out_h, int out_w) {
  const int PW = stem_fused_pitch(out_w);
  const long long tiles = (static_cast<long long>(out_h) * PW + kSTile - 1) / kSTile;
  return tiles * kSTile + 3ll * PW + 16 + 2 * kSTile;   // read slack: tiles of the pooled schedule may start past a tile boundary
}

int stem_rows(const uint8_t* img, const float* img_f32, int B, int IH, int IW, int crop_y, int crop_x, int H, int W, int RH, int RW,
              float rscale, const float* mean3, const float* std3, float post_scale, float post_shift, __nv_bfloat16* out,
              cudaStream_t stream) {
  const DeviceInfo* di = device_info();
  if (!di) return -2;
  DCR_REQUIRE(crop_y >= 0 && crop_x >= 0 && crop_y + H <= IH && crop_x + W <= IW, "stem_rows: crop outside image");
  if (rscale == 0.f) { RH = H; RW = W; }
  DCR_REQUIRE(RH >= 2 && RW >= 2 && RH % 2 == 0 && RW % 2 == 0, "stem_rows: network input size must be even (%d x %d)", RH, RW);
  if (B == 0) return 0;
  StemRowsParams p;
  p.img = img; p.img_f32 = img_f32; p.B = B; p.IH = IH; p.IW = IW; p.crop_y = crop_y; p.crop_x = crop_x; p.H = H; p.W = W;
  p.RH = RH; p.RW = RW; p.rscale = rscale;
  for (int c = 0; c < 3; ++c) { p.mean[c] = mean3[c]; p.std[c] = std3[c]; }
  p.post_scale = post_scale; p.post_shift = post_shift;
  const int OH = RH / 2, OW = RW / 2;
  p.PW = stem_fused_pitch(OW);
  p.rows = OH + 3;
  p.plane_stride = stem_fused_plane_units(OH, OW) * 8;
  p.img_stride = 2 * p.plane_stride;
  p.out = out;
  const long long total = 2ll * p.rows * p.PW * B;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(di->num_sms) * 16));
  if (img_f32) {
    if (rscale == 0.f) stem_rows_kernel<false, true><<<grid, 256, 0, stream>>>(p);
    else stem_rows_kernel<true, true><<<grid, 256, 0, stream>>>(p);
  } else {
    if (rscale == 0.f) stem_rows_kernel<false, false><<<grid, 256, 0, stream>>>(p);
    else stem_rows_kernel<true, false><<<grid, 256, 0, stream>>>(p);
  }
  count_launch();
  DCR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int stem_conv(const __nv_bfloat16* planes, int B, int OH, int OW, const __nv_bfloat16* weight, const float* scale, const float* bias,
              __nv_bfloat16* out, cudaStream_t stream, int pool) {
  const DeviceInfo* di = device_info();
  if (!di) return -2;
  DCR_REQUIRE(di->cc_major == 10, "stem_conv: this build targets sm_100a; device reports sm_%d%d", di->cc_major, di->cc_minor);
  if (B == 0) return 0;
  StemParams p;
  memset(&p, 0, sizeof(p));
  p.planes = planes; p.B = B; p.OH = OH; p.OW = OW;
  p.PW = stem_fused_pitch(OW);
  p.plane_stride = stem_fused_plane_units(OH, OW) * 8;
  p.img_stride = 2 * p.plane_stride;
  const int tiles_per_img = static_cast<int>((static_cast<long long>(OH) * p.PW + kSTile - 1) / kSTile);
  p.win_units = (kSTile + 3 * p.PW + 3 + 7) & ~7;
  p.scale = scale; p.bias = bias; p.out = out;
  p.OHp = (OH - 1) / 2 + 1;
  p.OWp = (OW - 1) / 2 + 1;
  if (pool) {
    const int parts = (p.OHp % 4 == 0) ? 4 : ((p.OHp % 2 == 0) ? 2 : 1);
    p.units_per_img = parts;
    p.prow_per_part = p.OHp / parts;
    p.part_rows = 2 * p.prow_per_part;
    p.tiles_per_unit = static_cast<int>((static_cast<long long>(p.part_rows + 1) * p.PW + kSTile - 1) / kSTile);
    // the last part's tiles may run past the image's last conv row: the planes carry zero slack for those reads
    const long long last_unit = static_cast<long long>(std::max(0, (parts - 1) * p.part_rows - 1)) * p.PW;
    DCR_REQUIRE(last_unit + static_cast<long long>(p.tiles_per_unit) * kSTile + 3ll * p.PW + 8 <= stem_fused_plane_units(OH, OW),
                "stem_conv: plane slack too small for the pooled schedule (%d x %d)", OH, OW);
  } else {
    p.units_per_img = tiles_per_img;
    p.tiles_per_unit = 1;
    p.prow_per_part = 0;
    p.part_rows = 0;
  }
  p.num_units = B * p.units_per_img;
  CUtensorMap tw;
  if (int rc = make_tmap_2d_bf16(&tw, weight, kSN, 256, 256, kSN, 64)) return rc;
  const size_t stage = (static_cast<size_t>(2) * p.win_units * 16 + 1023) & ~size_t(1023);
  const size_t out_bytes = pool ? ((static_cast<size_t>(8) * OW * 128 + 1023) & ~size_t(1023)) : 2 * 16384;
  const size_t fixed = 1024 + kWBytes + out_bytes + 512 + 256;
  DCR_REQUIRE(fixed + 2 * stage <= di->max_smem_optin, "stem_conv: image too wide for the window buffers (OW = %d)", OW);
  p.win_stages = static_cast<int>(std::min<size_t>(4, (di->max_smem_optin - fixed) / stage));
  const size_t smem = fixed + p.win_stages * stage;
  static bool attr_set[64][2] = {};
  const int grid = std::min(p.num_units, di->num_sms);
  if (pool) {
    if (!attr_set[di->device & 63][1]) {
      DCR_CUDA_CHECK(cudaFuncSetAttribute(stem_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(di->max_smem_optin)));
      attr_set[di->device & 63][1] = true;
    }
    stem_conv_kernel<true><<<grid, kSThreads, smem, stream>>>(tw, p);
  } else {
    if (!attr_set[di->device & 63][0]) {
      DCR_CUDA_CHECK(cudaFuncSetAttribute(stem_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(di->max_smem_optin)));
      attr_set[di->device & 63][0] = true;
    }
    stem_conv_kernel<false><<<grid, kSThreads, smem, stream>>>(tw, p);
  }
  count_launch();
  DCR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace dcr
