// conv3x3.cu -- drp_conv3x3: 3x3 convolution (padding 1) + bias + ReLU + {identity | 2x2 max-pool | 2x nearest upsample} as one
// implicit-GEMM kernel on the 5th-generation tensor cores (tcgen05, TF32 operands, fp32 accumulators in TMEM).
// Replaces the layers of the reference's OIDN-style U-Net denoiser, diffrp/rendering/denoiser.py:42-66 (Conv, relu, pool, upsample,
// concat) as used by UNet.forward (:117-173); see include/diffrp_b200.h for the tensor layout contract.
//
// GEMM view per CTA:  D[128 pixels, cout_pad] = sum over 9 taps x (cin/16) channel chunks of  A[128, 16] * B[cout_pad, 16]^T
//   * M tile = 8 rows x 16 columns of output pixels.  For tap (ky,kx) the A operand is the same 8x16 pixel box shifted by (ky-1,kx-1):
//     ONE 3-D TMA box load (16 channels, 16 x, 8 y) per k-step; pixels outside the image are zero-filled by TMA (= padding 1).
//   * B operand = 16 consecutive k of every output channel: one 2-D TMA box (16, cout_pad) of the [cout_pad][9*cin] weight matrix.
//   * both tiles are K-major rows of 64 B -> SWIZZLE_64B on the TMA side and in the UMMA shared-memory descriptors;
//     two tcgen05.mma.kind::tf32 (K = 8) per k-step, issued by one thread; tcgen05.commit releases the stage / signals the epilogue.
//   * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue (TMEM lane quadrant = warp % 4):
//     tcgen05.ld 32x32b.x16 -> bias, ReLU -> pool via two shuffles (a 2x2 window lives in one warp) or 4-way replicated store.
//   * STAGES-deep mbarrier ring between producer and MMA; several CTAs per SM overlap one tile's epilogue with another's main loop.
#include <cuda.h>
#include <cstdint>
#include <mutex>
#include <algorithm>
#include <cstdlib>
#include <string>
#include "internal.h"

namespace {

constexpr int TILE_W = 16, TILE_H = 8, BM = TILE_W * TILE_H;  // 128 output pixels per CTA = UMMA M
constexpr int OG = 16;                                         // output channels per epilogue group: 64-byte staging rows (SWIZZLE_64B)
// input-channel chunk per k-step (template parameter KCH): 16 fp32 = 64 B rows (SWIZZLE_64B) or 32 fp32 = 128 B rows (SWIZZLE_128B)
constexpr int MAX_STAGES = 8;
constexpr int NUM_THREADS = 192;
constexpr uint32_t SPIN_LIMIT = 1u << 28;                      // a wedged pipeline traps instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
        if (spin > SPIN_LIMIT) __trap();
    }
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// UMMA shared-memory descriptor, K-major, rows of ROW_BYTES = 64 (SWIZZLE_64B) or 128 (SWIZZLE_128B): 8-row groups 8*ROW_BYTES apart (SBO);
// version 1 (sm_100)
template <int ROW_BYTES>
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);           // start address, bits [0,14)
    d |= (uint64_t)((8u * ROW_BYTES) >> 4) << 32;            // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                  // descriptor version
    d |= (uint64_t)(ROW_BYTES == 128 ? 2 : 4) << 61;         // layout type: SWIZZLE_128B = 2, SWIZZLE_64B = 4
    return d;
}
// instruction descriptor: D fp32, A/B TF32, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct ConvArgs {
    const float* bias;
    float* out;
    int height, width, k_steps, chunks;  // chunks = cin / 16, k_steps = 3 * chunks (one per filter column and channel chunk)
    int cin, cout_pad, cout_store, out_stride, out_offset, mode, relu, round_tf32;
    int tmem_cols, stages;
};

__device__ __forceinline__ float round_tf32(float v) {  // round-to-nearest on the 10-bit TF32 mantissa: the MMA then truncates nothing
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
// Epilogue of one tile for one warp (TMEM lane quadrant `quad` = 2 rows x 16 pixels of every 8-row half): TMEM -> registers (bias, ReLU,
// TF32 rounding, 2x2 max by shuffles) -> this warp's private 2 KB staging buffer (64 B rows, SWIZZLE_64B) -> one TMA tensor store per
// 16-channel group.  No CTA-wide barrier: the four epilogue warps run independently; the store's tensor map does the addressing
// (channel slice + pixel stride, clipping at the image border and at cout_store, and the 2x replication of the upsampling layers).
template <int TR>
__device__ __forceinline__ void epilogue_tile(const ConvArgs& a, const CUtensorMap* map_out, uint32_t tacc, int quad, int lane, int x0, int y0,
                                              uint8_t* warp_stage, int& buf) {
    constexpr int HALVES = TR / 8;
    constexpr int WARP_OUT_BYTES = 32 * OG * 4;   // 32 pixels x 16 channels
    const int tx = lane & 15;
    const uint64_t mo = reinterpret_cast<uint64_t>(map_out);
    for (int c0 = 0; c0 < a.cout_store; c0 += 16) {
#pragma unroll
        for (int half = 0; half < HALVES; ++half, buf ^= 1) {
            uint8_t* stage_out = warp_stage + buf * WARP_OUT_BYTES;
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store that last used this buffer has read it
            __syncwarp();
            float v[16];
            tmem_ld16(tacc + (uint32_t)(half * a.cout_pad + c0), v);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                v[i] += __ldg(a.bias + c0 + i);
                if (a.relu) v[i] = fmaxf(v[i], 0.0f);
                if (a.round_tf32) v[i] = round_tf32(v[i]);
            }
            int row = lane;
            bool writer = true;
            if (a.mode == DRP_CONV_POOL2) {       // the warp's two tile rows: lanes l, l^1, l^16, l^17 form a 2x2 window
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    v[i] = fmaxf(v[i], __shfl_xor_sync(0xffffffffu, v[i], 1));
                    v[i] = fmaxf(v[i], __shfl_xor_sync(0xffffffffu, v[i], 16));
                }
                writer = (lane & 17) == 0;
                row = tx >> 1;
            }
            if (writer) {
                float4* dst = reinterpret_cast<float4*>(stage_out + row * (OG * 4));
                const int sw = (row >> 1) & 3;    // SWIZZLE_64B: 16-byte chunk index ^= address bits [7,9)
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[j ^ sw] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA engine
            __syncwarp();
            if (lane == 0) {
                const uint32_t src = smem_u32(stage_out);
                const int yy = y0 + half * 8 + quad * 2;   // first of this warp's two rows
                if (a.mode == DRP_CONV_PLAIN) {
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                                 ::"l"(mo), "r"(src), "r"(c0), "r"(x0), "r"(yy) : "memory");
                } else if (a.mode == DRP_CONV_POOL2) {
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                                 ::"l"(mo), "r"(src), "r"(c0), "r"(x0 >> 1), "r"(yy >> 1) : "memory");
                } else {
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                                     ::"l"(mo), "r"(src), "r"(c0), "r"(r & 1), "r"(x0), "r"(r >> 1), "r"(yy) : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
}

// TR = output rows per CTA (8 or 16): the M = 128 accumulator(s) cover 8 rows x 16 columns each.
// Per k-step (filter column kx, 16-channel chunk) ONE activation box of TR+2 rows is loaded; the three filter rows are the same
// shared-memory tile read at row offsets 0 / 1 / 2 (a row of 16 pixels = 1024 B, a multiple of the swizzle period), so every
// activation element crosses L2 -> SM 3 (TR+2)/TR times per layer instead of 9, and the weights once per TR*16 pixels.
template <int TR, int KCH>
__global__ void __launch_bounds__(NUM_THREADS) k_conv3x3_tf32(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                              const __grid_constant__ CUtensorMap map_out, const ConvArgs a) {
    constexpr int HALVES = TR / 8;
    constexpr int A_BYTES = (TR + 2) * TILE_W * KCH * 4;    // (TR+2) or 2(TR+2) KB, 1024-aligned
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int b_tap_bytes = a.cout_pad * KCH * 4;
    const int b_stage_bytes = 3 * b_tap_bytes;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + a.stages * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_b + a.stages * b_stage_bytes);
    uint64_t* empty = full + MAX_STAGES;
    uint64_t* acc_ready = empty + MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_ready + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * TR;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_out)) : "memory");
        for (int s = 0; s < a.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: one warp allocates (and later frees) the accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)a.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer ----
            const uint32_t stage_bytes = (uint32_t)(A_BYTES + b_stage_bytes);
            int s = 0; uint32_t phase = 0;
            for (int kb = 0; kb < a.k_steps; ++kb) {
                mbar_wait(&empty[s], phase ^ 1u);
                mbar_expect_tx(&full[s], stage_bytes);
                const int kx = kb / a.chunks, chunk = kb - kx * a.chunks;
                tma_load_3d(&map_a, &full[s], smem_a + s * A_BYTES, chunk * KCH, x0 + kx - 1, y0 - 1);
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
                    tma_load_2d(&map_b, &full[s], smem_b + s * b_stage_bytes + ky * b_tap_bytes, (ky * 3 + kx) * a.cin + chunk * KCH, 0);
                if (++s == a.stages) { s = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer ----
            const uint32_t idesc = umma_idesc_tf32(a.cout_pad);
            int s = 0; uint32_t phase = 0;
            for (int kb = 0; kb < a.k_steps; ++kb) {
                mbar_wait(&full[s], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t pa = smem_u32(smem_a + s * A_BYTES), pb = smem_u32(smem_b + s * b_stage_bytes);
                // descriptors advance in their 14-bit start-address field (16-byte units): A by whole tile rows (1024 B) and 32 B per UMMA K
                const uint64_t da0 = umma_desc<KCH * 4>(pa), db0 = umma_desc<KCH * 4>(pb);
                const uint32_t b_tap16 = (uint32_t)b_tap_bytes >> 4;
#pragma unroll
                for (int half = 0; half < HALVES; ++half)
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int k = 0; k < KCH / 8; ++k)  // UMMA K = 8 TF32 = 32 B along the row
                            umma_tf32(tmem_base + (uint32_t)(half * a.cout_pad), da0 + (uint64_t)((half * 8 + ky) * (TILE_W * KCH * 4 / 16) + k * 2),
                                      db0 + (uint64_t)(ky * b_tap16 + k * 2), idesc, (kb | ky | k) != 0);
                umma_commit(&empty[s]);           // frees the stage when these MMAs have read it
                if (++s == a.stages) { s = 0; phase ^= 1u; }
            }
            umma_commit(acc_ready);               // accumulators complete
        }
    } else {
        // ---- epilogue: warps 2..5, TMEM lane quadrant = warp % 4; staging aliases the (drained) pipeline stages ----
        const int quad = warp & 3;
        mbar_wait(acc_ready, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        int buf = 0;
        epilogue_tile<TR>(a, &map_out, tmem_base + ((uint32_t)(quad * 32) << 16), quad, lane, x0, y0, smem_a + (warp - 2) * (2 * 32 * OG * 4), buf);
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before the CTA retires its shared memory
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols) : "memory");
    }
}

// ---- persistent variant: one CTA per SM loops over tiles; the accumulator is double-buffered in TMEM so that the epilogue of tile i
// (TMEM -> registers -> smem -> TMA store) overlaps the TMA loads and MMAs of tile i+1; the smem stage ring runs across tile boundaries.
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int TR, int KCH>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_conv3x3_tf32_persistent(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                                            const __grid_constant__ CUtensorMap map_out, const ConvArgs a) {
    constexpr int HALVES = TR / 8;
    constexpr int A_BYTES = (TR + 2) * TILE_W * KCH * 4;
    constexpr int OUT_BYTES = 4 * 32 * OG * 4;             // per-warp double-buffered staging: 4 warps x 2 x 2 KB = 2 x OUT_BYTES
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int b_tap_bytes = a.cout_pad * KCH * 4;
    const int b_stage_bytes = 3 * b_tap_bytes;
    uint8_t* smem_out = smem;                                   // 2 x OUT_BYTES epilogue staging
    uint8_t* smem_a = smem + 2 * OUT_BYTES;
    uint8_t* smem_b = smem_a + a.stages * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_b + a.stages * b_stage_bytes);
    uint64_t* empty = full + MAX_STAGES;
    uint64_t* acc_full = empty + MAX_STAGES;                    // [2]
    uint64_t* acc_empty = acc_full + 2;                         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (a.width + TILE_W - 1) / TILE_W, tiles_y = (a.height + TR - 1) / TR;
    const int n_tiles = tiles_x * tiles_y;
    const uint32_t acc_stride = (uint32_t)(HALVES * a.cout_pad);  // TMEM columns per accumulator buffer

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_out)) : "memory");
        for (int s = 0; s < a.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)a.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer ----
            const uint32_t stage_bytes = (uint32_t)(A_BYTES + b_stage_bytes);
            int s = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int by = tile / tiles_x, bx = tile - by * tiles_x;
                const int x0 = bx * TILE_W, y0 = by * TR;
                for (int kb = 0; kb < a.k_steps; ++kb) {
                    mbar_wait(&empty[s], phase ^ 1u);
                    mbar_expect_tx(&full[s], stage_bytes);
                    const int kx = kb / a.chunks, chunk = kb - kx * a.chunks;
                    tma_load_3d(&map_a, &full[s], smem_a + s * A_BYTES, chunk * KCH, x0 + kx - 1, y0 - 1);
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
                        tma_load_2d(&map_b, &full[s], smem_b + s * b_stage_bytes + ky * b_tap_bytes, (ky * 3 + kx) * a.cin + chunk * KCH, 0);
                    if (++s == a.stages) { s = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer ----
            const uint32_t idesc = umma_idesc_tf32(a.cout_pad);
            int s = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const int ab = it & 1;
                mbar_wait(&acc_empty[ab], (((uint32_t)it >> 1) & 1u) ^ 1u);   // the epilogue has drained this accumulator buffer
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_base + (uint32_t)ab * acc_stride;
                for (int kb = 0; kb < a.k_steps; ++kb) {
                    mbar_wait(&full[s], phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t pa = smem_u32(smem_a + s * A_BYTES), pb = smem_u32(smem_b + s * b_stage_bytes);
                    const uint64_t da0 = umma_desc<KCH * 4>(pa), db0 = umma_desc<KCH * 4>(pb);
                    const uint32_t b_tap16 = (uint32_t)b_tap_bytes >> 4;
#pragma unroll
                    for (int half = 0; half < HALVES; ++half)
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                            for (int k = 0; k < KCH / 8; ++k)
                                umma_tf32(tacc + (uint32_t)(half * a.cout_pad), da0 + (uint64_t)((half * 8 + ky) * (TILE_W * KCH * 4 / 16) + k * 2),
                                          db0 + (uint64_t)(ky * b_tap16 + k * 2), idesc, (kb | ky | k) != 0);
                    umma_commit(&empty[s]);
                    if (++s == a.stages) { s = 0; phase ^= 1u; }
                }
                umma_commit(&acc_full[ab]);
            }
        }
    } else {
        // ---- epilogue warps 2..5 ----
        const int quad = warp & 3;
        uint8_t* warp_stage = smem_out + (warp - 2) * (2 * 32 * OG * 4);
        int buf = 0, it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int by = tile / tiles_x, bx = tile - by * tiles_x;
            const int ab = it & 1;
            mbar_wait(&acc_full[ab], ((uint32_t)it >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            epilogue_tile<TR>(a, &map_out, tmem_base + (uint32_t)ab * acc_stride + ((uint32_t)(quad * 32) << 16), quad, lane, bx * TILE_W, by * TR, warp_stage, buf);
            // every TMEM read of this tile has completed (tcgen05.wait::ld in tmem_ld16): hand the accumulator buffer back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&acc_empty[ab]);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols) : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// per-device one-time setup: opt in to > 48 KB of dynamic shared memory for every kernel variant, remember the SM count
struct ConvDeviceState { bool ready = false; cudaError_t err = cudaSuccess; int sm_count = 148; };
static ConvDeviceState* conv_device_state() {
    static std::mutex mu;
    static ConvDeviceState states[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    std::lock_guard<std::mutex> lk(mu);
    ConvDeviceState& st = states[dev];
    if (!st.ready) {
        const int limit = 220 * 1024;
        const void* kernels[8] = {(const void*)k_conv3x3_tf32<8, 16>, (const void*)k_conv3x3_tf32<16, 16>, (const void*)k_conv3x3_tf32<8, 32>,
                                  (const void*)k_conv3x3_tf32<16, 32>, (const void*)k_conv3x3_tf32_persistent<8, 16>,
                                  (const void*)k_conv3x3_tf32_persistent<16, 16>, (const void*)k_conv3x3_tf32_persistent<8, 32>,
                                  (const void*)k_conv3x3_tf32_persistent<16, 32>};
        for (int k = 0; k < 8 && st.err == cudaSuccess; ++k) st.err = cudaFuncSetAttribute(kernels[k], cudaFuncAttributeMaxDynamicSharedMemorySize, limit);
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) st.sm_count = prop.multiProcessorCount;
        st.ready = true;
    }
    return &st;
}

}  // namespace

// Tile / stage overrides for tools/tune_conv.py: compiled in only with -DDRP_EXPERIMENTAL (DRP_NVCC_EXTRA); the shipped library reads no
// environment variable.
static inline const char* tune_env(const char* name) {
#ifdef DRP_EXPERIMENTAL
    return getenv(name);
#else
    (void)name;
    return nullptr;
#endif
}

extern "C" int drp_conv3x3(const drp_conv3x3_params_t* pp, void* stream) {
    if (!pp) { drp_set_error("drp_conv3x3: params is NULL"); return DRP_ERR_INVALID; }
    const drp_conv3x3_params_t p = *pp;
    if (p.height <= 0 || p.width <= 0) { drp_set_error("drp_conv3x3: empty image"); return DRP_ERR_INVALID; }
    if (!p.in || !p.weight || !p.bias || !p.out) { drp_set_error("drp_conv3x3: NULL tensor"); return DRP_ERR_INVALID; }
    if (p.cin < 16 || p.cin % 16 || p.in_stride % 16 || p.in_offset % 16 || p.in_offset + p.cin > p.in_stride) {
        drp_set_error("drp_conv3x3: cin / in_stride / in_offset must be multiples of 16 with in_offset + cin <= in_stride"); return DRP_ERR_INVALID; }
    if (p.cout_pad < 16 || p.cout_pad > 256 || p.cout_pad % 16 || p.cout_store < 4 || p.cout_store % 4 || p.cout_store > p.cout_pad) {
        drp_set_error("drp_conv3x3: cout_pad must be a multiple of 16 in [16, 256], cout_store a multiple of 4 (TMA stores 16-byte units) <= cout_pad"); return DRP_ERR_INVALID; }
    if (p.out_stride % 4 || p.out_offset % 4 || p.out_offset + p.cout_store > p.out_stride) {
        drp_set_error("drp_conv3x3: out_stride / out_offset must be multiples of 4 with out_offset + cout_store <= out_stride"); return DRP_ERR_INVALID; }
    if (p.mode < DRP_CONV_PLAIN || p.mode > DRP_CONV_UPSAMPLE2) { drp_set_error("drp_conv3x3: unknown mode"); return DRP_ERR_INVALID; }
    if (p.mode == DRP_CONV_POOL2 && ((p.height | p.width) & 1)) { drp_set_error("drp_conv3x3: pooling needs even height and width"); return DRP_ERR_INVALID; }
    if ((reinterpret_cast<uintptr_t>(p.in) | reinterpret_cast<uintptr_t>(p.weight) | reinterpret_cast<uintptr_t>(p.out)) & 15) {
        drp_set_error("drp_conv3x3: tensors must be 16-byte aligned"); return DRP_ERR_INVALID; }
    EncodeTiledFn encode = encode_tiled();
    if (!encode) { drp_set_error("drp_conv3x3: cuTensorMapEncodeTiled is unavailable in this driver"); return DRP_ERR_CUDA; }

    // rows per CTA: 16 halves the weight traffic and the halo overhead, 8 keeps every SM busy on the small (deep) levels
    const int64_t tiles16 = (int64_t)((p.width + TILE_W - 1) / TILE_W) * ((p.height + 15) / 16);
    const bool wide_ok = p.cout_pad <= 128 && p.cout_pad > 16;
    const bool persist16 = wide_ok && tiles16 > 148 && tiles16 <= 320;              // 1-2 waves of 16-row tiles: persistent kernel (see below)
    int tr = wide_ok && (tiles16 >= 2 * 148 || persist16) ? 16 : 8;                  // per-layer sweep: tools/tune_conv.py
    if (const char* e = tune_env("DRP_CONV_ROWS")) { const int v = atoi(e); if (v == 8 || (v == 16 && p.cout_pad <= 128)) tr = v; }
    // input channels per k-step: 16 (64-byte rows, SWIZZLE_64B) or 32 (128-byte rows, SWIZZLE_128B: half the k-steps, stages twice as big).
    // Per-layer A/B (tools/tune_conv.py, TUNE_KC=1): 32 only pays on the latency-bound sub-wave grids of the deep levels (20.8 -> 18.4 us);
    // on the wide layers the bigger stages cost co-resident CTAs (82 -> 108 us, 90 -> 125 us), so 16 stays the default there.
    const int64_t tiles8 = (int64_t)((p.width + TILE_W - 1) / TILE_W) * ((p.height + 7) / 8);
    int kch = p.cin % 32 == 0 && tiles8 <= 148 ? 32 : 16;
    if (const char* e = tune_env("DRP_CONV_KC")) { const int v = atoi(e); if (v == 16 || (v == 32 && p.cin % 32 == 0)) kch = v; }
    const CUtensorMapSwizzle in_swizzle = kch == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUtensorMap map_a, map_b;
    {   // activations: (C, W, H) fp32 view of the channel slice, box (kch, 16, tr + 2), zero fill outside = padding 1
        const cuuint64_t dims[3] = {(cuuint64_t)p.cin, (cuuint64_t)p.width, (cuuint64_t)p.height};
        const cuuint64_t strides[2] = {(cuuint64_t)p.in_stride * 4, (cuuint64_t)p.in_stride * 4 * (cuuint64_t)p.width};
        const cuuint32_t box[3] = {(cuuint32_t)kch, TILE_W, (cuuint32_t)(tr + 2)}, estr[3] = {1, 1, 1};
        CUresult r = encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p.in + p.in_offset), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, in_swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { drp_set_error("drp_conv3x3: cuTensorMapEncodeTiled(activations) failed: " + std::to_string((int)r)); return DRP_ERR_CUDA; }
    }
    {   // weights: (K = 9*cin, cout_pad), box (16, cout_pad)
        const cuuint64_t dims[2] = {(cuuint64_t)9 * p.cin, (cuuint64_t)p.cout_pad};
        const cuuint64_t strides[1] = {(cuuint64_t)9 * p.cin * 4};
        const cuuint32_t box[2] = {(cuuint32_t)kch, (cuuint32_t)p.cout_pad}, estr[2] = {1, 1};
        CUresult r = encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p.weight), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, in_swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { drp_set_error("drp_conv3x3: cuTensorMapEncodeTiled(weights) failed: " + std::to_string((int)r)); return DRP_ERR_CUDA; }
    }
    CUtensorMap map_out;
    {   // output: channel slice [out_offset, out_offset + cout_store) of the NHWC buffer; the store clips at the image border and at cout_store
        float* base = p.out + p.out_offset;
        const cuuint64_t ps = (cuuint64_t)p.out_stride * 4;   // pixel stride in bytes
        CUresult r;
        if (p.mode == DRP_CONV_UPSAMPLE2) {                    // (c, b, x, a, y): output pixel (2y + a, 2x + b)
            const cuuint64_t dims[5] = {(cuuint64_t)p.cout_store, 2, (cuuint64_t)p.width, 2, (cuuint64_t)p.height};
            const cuuint64_t strides[4] = {ps, 2 * ps, 2 * (cuuint64_t)p.width * ps, 4 * (cuuint64_t)p.width * ps};
            const cuuint32_t box[5] = {OG, 1, TILE_W, 1, 2}, estr[5] = {1, 1, 1, 1, 1};   // one warp's 2 rows x 16 pixels per store
            r = encode(&map_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {
            const int sh = p.mode == DRP_CONV_POOL2 ? 1 : 0;
            const cuuint64_t ow = (cuuint64_t)(p.width >> sh), oh = (cuuint64_t)(p.height >> sh);
            const cuuint64_t dims[3] = {(cuuint64_t)p.cout_store, ow, oh};
            const cuuint64_t strides[2] = {ps, ow * ps};
            const cuuint32_t box[3] = {OG, (cuuint32_t)(TILE_W >> sh), (cuuint32_t)(2 >> sh)}, estr[3] = {1, 1, 1};   // one warp's 2 rows x 16 pixels (pooled: 1 x 8)
            r = encode(&map_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (r != CUDA_SUCCESS) { drp_set_error("drp_conv3x3: cuTensorMapEncodeTiled(output) failed: " + std::to_string((int)r)); return DRP_ERR_CUDA; }
    }
    ConvDeviceState* dstate = conv_device_state();
    DRP_CUDA_CHECK(dstate->err);
    ConvArgs a;
    a.bias = p.bias; a.out = p.out; a.height = p.height; a.width = p.width;
    a.chunks = p.cin / kch; a.k_steps = 3 * a.chunks; a.cin = p.cin;
    a.cout_pad = p.cout_pad; a.cout_store = p.cout_store; a.out_stride = p.out_stride; a.out_offset = p.out_offset; a.mode = p.mode; a.relu = p.relu;
    a.round_tf32 = p.round_tf32;
    const int acc_cols = (tr / 8) * p.cout_pad;
    a.tmem_cols = acc_cols <= 32 ? 32 : acc_cols <= 64 ? 64 : acc_cols <= 128 ? 128 : acc_cols <= 256 ? 256 : 512;
    const size_t stage_bytes = (size_t)(tr + 2) * TILE_W * kch * 4 + 3 * (size_t)p.cout_pad * kch * 4;
    // shared-memory budget per CTA (tools/tune_conv.py): ~48 KB keeps 3-4 CTAs per SM so that one tile's epilogue overlaps the others'
    // main loops; grids below two waves are latency-bound and want every stage they can get; the 16-channel output layer is store-bound
    const int64_t n_tiles = (int64_t)((p.width + TILE_W - 1) / TILE_W) * ((p.height + tr - 1) / tr);
    size_t budget = n_tiles < 2 * 148 ? 100 * 1024 : p.cout_pad <= 16 ? 24 * 1024 : 48 * 1024;
    if (const char* e = tune_env("DRP_CONV_SMEM_KB")) budget = (size_t)atoi(e) * 1024;
    a.stages = (int)std::min<size_t>(MAX_STAGES, std::max<size_t>(2, budget / stage_bytes));
    a.stages = std::min(a.stages, std::max(2, a.k_steps));
    // persistent variant (one CTA per SM, double-buffered TMEM accumulator).  Per-layer A/B (tools/tune_conv.py, TUNE_PERSISTENT=1): it wins
    // only for 1-2 waves of tiles (U-Net levels at 1/4 resolution: 51 -> 44 us, 52 -> 42 us); with many tiles per SM a single epilogue warp
    // group per SM is slower than 3-4 co-resident one-tile CTAs (full-resolution layers: 137 -> 210 us, 62 -> 142 us), so those keep the
    // one-tile kernel.
    bool persistent = tr == 16 && n_tiles > 148 && n_tiles <= 320 && 2 * acc_cols <= 512;
    if (const char* e = tune_env("DRP_CONV_PERSISTENT")) persistent = atoi(e) != 0 && 2 * acc_cols <= 512;
    if (persistent) {
        const int pacc = 2 * acc_cols;
        a.tmem_cols = pacc <= 32 ? 32 : pacc <= 64 ? 64 : pacc <= 128 ? 128 : pacc <= 256 ? 256 : 512;
        size_t pbudget = 190 * 1024 - 16 * 1024;
        if (const char* e = tune_env("DRP_CONV_SMEM_KB")) pbudget = (size_t)atoi(e) * 1024;
        a.stages = (int)std::min<size_t>(MAX_STAGES, std::max<size_t>(2, pbudget / stage_bytes));
        const size_t psmem = 1024 + 16 * 1024 + (size_t)a.stages * stage_bytes + (2 * MAX_STAGES + 4) * sizeof(uint64_t) + 16;
        if (psmem <= 220 * 1024) {
            const unsigned grid = (unsigned)std::min<int64_t>(n_tiles, dstate->sm_count);
            cudaStream_t st = (cudaStream_t)stream;
            if (tr == 16 && kch == 32) k_conv3x3_tf32_persistent<16, 32><<<grid, NUM_THREADS, psmem, st>>>(map_a, map_b, map_out, a);
            else if (tr == 16) k_conv3x3_tf32_persistent<16, 16><<<grid, NUM_THREADS, psmem, st>>>(map_a, map_b, map_out, a);
            else if (kch == 32) k_conv3x3_tf32_persistent<8, 32><<<grid, NUM_THREADS, psmem, st>>>(map_a, map_b, map_out, a);
            else k_conv3x3_tf32_persistent<8, 16><<<grid, NUM_THREADS, psmem, st>>>(map_a, map_b, map_out, a);
            DRP_CUDA_CHECK(cudaGetLastError());
            return DRP_OK;
        }
    }
    const size_t smem = 1024 + (size_t)a.stages * stage_bytes + (2 * MAX_STAGES + 1) * sizeof(uint64_t) + 16;
    if (smem > 220 * 1024) { drp_set_error("drp_conv3x3: tile does not fit in shared memory"); return DRP_ERR_INVALID; }
    const dim3 grid((unsigned)((p.width + TILE_W - 1) / TILE_W), (unsigned)((p.height + tr - 1) / tr));
    cudaStream_t st = (cudaStream_t)stream;
    if (tr == 16 && kch == 32) k_conv3x3_tf32<16, 32><<<grid, NUM_THREADS, smem, st>>>(map_a, map_b, map_out, a);
    else if (tr == 16) k_conv3x3_tf32<16, 16><<<grid, NUM_THREADS, smem, st>>>(map_a, map_b, map_out, a);
    else if (kch == 32) k_conv3x3_tf32<8, 32><<<grid, NUM_THREADS, smem, st>>>(map_a, map_b, map_out, a);
    else k_conv3x3_tf32<8, 16><<<grid, NUM_THREADS, smem, st>>>(map_a, map_b, map_out, a);
    DRP_CUDA_CHECK(cudaGetLastError());
    return DRP_OK;
}

// ---- network input / output transforms of run_denoiser (diffrp/rendering/denoiser.py:24-35) ------------------------------------
// PU transfer function, diffrp/utils/colors.py:5-31
namespace {
constexpr float PU_A = 1.41283765e+03f, PU_B = 1.64593172e+00f, PU_C = 4.31384981e-01f, PU_D = -2.94139609e-03f, PU_E = 1.92653254e-01f,
                PU_F = 6.26026094e-03f, PU_G = 9.98620152e-01f, PU_Y0 = 1.57945760e-06f, PU_Y1 = 3.22087631e-02f, PU_X0 = 2.23151711e-03f,
                PU_X1 = 3.70974749e-01f;
__host__ __device__ inline float linear_to_pu(float y) {
    return y <= PU_Y0 ? PU_A * y : (y <= PU_Y1 ? PU_B * powf(y, PU_C) + PU_D : PU_E * logf(y + PU_F) + PU_G);
}
__host__ __device__ inline float pu_to_linear(float x) {
    return x <= PU_X0 ? x / PU_A : (x <= PU_X1 ? powf((x - PU_D) / PU_B, 1.0f / PU_C) : expf((x - PU_G) / PU_E) - PU_F);
}
__device__ __forceinline__ int reflect_index(int i, int n) {  // nn.ReflectionPad2d: mirror without repeating the border sample
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return min(max(i, 0), n - 1);
}

__global__ void __launch_bounds__(256) k_denoise_pack(const float* __restrict__ hdr, const float* __restrict__ albedo, const float* __restrict__ normal,
                                                      int h, int w, float* __restrict__ dst, int H, int W, int stride, int offset, float pu_scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)H * W) return;
    const int Y = (int)(i / W), X = (int)(i - (int64_t)Y * W);
    const int top = (H - h) / 2, left = (W - w) / 2;          // dh // 2, dw // 2
    const int64_t s = ((int64_t)reflect_index(Y - top, h) * w + reflect_index(X - left, w)) * 3;
    float v[16];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        v[c] = linear_to_pu(__ldg(hdr + s + c)) * pu_scale;
        v[3 + c] = __ldg(albedo + s + c);
        v[6 + c] = __ldg(normal + s + c) * 0.5f + 0.5f;
    }
#pragma unroll
    for (int c = 0; c < 9; ++c) v[c] = round_tf32(v[c]);      // network input: convolution operand only
#pragma unroll
    for (int c = 9; c < 16; ++c) v[c] = 0.0f;
    float4* o = reinterpret_cast<float4*>(dst + i * stride + offset);
#pragma unroll
    for (int q = 0; q < 4; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

__global__ void __launch_bounds__(256) k_denoise_unpack(const float* __restrict__ src, int H, int W, int stride, float* __restrict__ out, int h, int w,
                                                        float inv_pu_scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)h * w) return;
    const int y = (int)(i / w), x = (int)(i - (int64_t)y * w);
    const int top = (H - h) / 2, left = (W - w) / 2;
    const float* s = src + ((int64_t)(y + top) * W + (x + left)) * stride;
#pragma unroll
    for (int c = 0; c < 3; ++c) out[3 * i + c] = pu_to_linear(__ldg(s + c) * inv_pu_scale);
}
}  // namespace

extern "C" int drp_denoise_pack(const float* hdr, const float* albedo_srgb, const float* normal, int32_t height, int32_t width, float* dst,
                                int32_t padded_height, int32_t padded_width, int32_t dst_stride, int32_t dst_offset, void* stream) {
    if (!hdr || !albedo_srgb || !normal || !dst || height < 1 || width < 1 || padded_height < height || padded_width < width ||
        dst_stride % 4 || dst_offset % 4 || dst_offset + 16 > dst_stride) {
        drp_set_error("drp_denoise_pack: invalid argument"); return DRP_ERR_INVALID; }
    if ((padded_height - height + 1) / 2 >= height || (padded_width - width + 1) / 2 >= width) {
        drp_set_error("drp_denoise_pack: reflection padding must be smaller than the image"); return DRP_ERR_INVALID; }
    const float pu_scale = 1.0f / linear_to_pu(65504.0f);
    const int64_t n = (int64_t)padded_height * padded_width;
    k_denoise_pack<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(hdr, albedo_srgb, normal, height, width, dst, padded_height, padded_width,
                                                                                   dst_stride, dst_offset, pu_scale);
    DRP_CUDA_CHECK(cudaGetLastError());
    return DRP_OK;
}

extern "C" int drp_denoise_unpack(const float* src, int32_t padded_height, int32_t padded_width, int32_t src_stride, float* out, int32_t height,
                                  int32_t width, void* stream) {
    if (!src || !out || height < 1 || width < 1 || padded_height < height || padded_width < width || src_stride < 3) {
        drp_set_error("drp_denoise_unpack: invalid argument"); return DRP_ERR_INVALID; }
    const float inv = linear_to_pu(65504.0f);
    const int64_t n = (int64_t)height * width;
    k_denoise_unpack<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, padded_height, padded_width, src_stride, out, height, width, inv);
    DRP_CUDA_CHECK(cudaGetLastError());
    return DRP_OK;
}
