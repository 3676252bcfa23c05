// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA -> smem ring -> tcgen05.mma -> TMEM -> fused epilogue.
//
//   C[M,N] = A[M,K] * B[N,K]^T  (A, B bf16 K-major; f32 accumulate in TMEM)
//
// One CTA per SM loops over 128 x BN output tiles (n fastest so the weight matrix stays L2-resident while an
// A panel is swept).  Roles: warp 0 = TMA producer, warp 1 = MMA issuer (convergent, one elected lane per instruction), warps 2..9 =
// epilogue (two warps per TMEM lane quarter, each owning half of the tile's columns).  The accumulator is
// double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the main loop of tile i+1.
//
// The A operand's K index may wrap onto following rows ("a_cols" < K): that is how the encoder's two Conv1d
// layers (k=3) run as implicit GEMMs over a zero-padded time-major activation without materialising im2col
// (CT2 ref: layers::WhisperEncoder conv1/conv2).  Epilogues fuse bias, exact-erf GELU,
// the residual add (f16 residual stream) and the positional-embedding add and remap GEMM rows to output
// rows so padded rows are never written.  The epilogue goes straight from the accumulator row a thread owns to global
// memory in 32-byte sectors (no shared-memory staging: shared memory belongs to the MMA operands).
#include <cuda_fp16.h>

#include <cstdlib>

#include "gemm.h"
#include "ptx.cuh"

namespace aries {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                       // 64 bf16 = 128 B = one SWIZZLE_128B atom row
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + kEpiWarps * 32;
constexpr bool kDefaultPair = true;          // CTA-pair kernel for BN = 256 (ARIES_GEMM_PAIR overrides)

constexpr int kBBoxRows = 128;               // rows of one B TMA box (the tensor maps are built with this box)

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, the two SMs of a TPC) per 256 x BN tile with
// tcgen05 cta_group::2 -- each CTA stages its own 128 rows of A and HALF of the B tile, so the shared-memory operand
// traffic per SM (the limiter of the 1-CTA 128 x 256 tile: 48 KB per 64-deep stage) drops by a third and two more
// stages fit.
template <int BN, int CG>
struct Cfg {
    static_assert(CG == 1 || (CG == 2 && BN == 256), "the CTA-pair kernel is built for BN = 256 only");
    static constexpr int kStages = (BN == 256 && CG == 1) ? 4 : (CG == 2 ? 7 : 6);   // no epilogue staging buffer any more
    static constexpr int kABytes = BM * BK * 2;
    static constexpr int kBBytes = (BN / CG) * BK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarBytes = 192;                              // 2 * stages + 4 mbarriers, the TMEM slot
    static constexpr int kVecBytes = 2 * BN * 4;                       // the tile's per-column vectors (bias | c1), f32
    static constexpr int kUsedBytes = kStages * kStageBytes + kBarBytes + kVecBytes;
    // + slack for aligning the operand ring to 1024 B (the dynamic segment is 1024-aligned in practice; the kernel traps
    // if the slack does not suffice rather than overrunning): everything that is left below the 227 KB limit
    static constexpr int kSmemBytes = (kUsedBytes + 1024 <= 232448) ? kUsedBytes + 1024 : 232448;
    static_assert(kUsedBytes + 512 <= 232448, "shared memory budget");
    static constexpr uint32_t kTmemCols = 2 * BN;
};

template <int BN, int EPI, int CG>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const GemmParams p) {
    using C = Cfg<BN, CG>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
    float* vec = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes + C::kBarBytes);   // [2][BN]: bias | c1 of the tile
    if ((base - smem_u32(smem_raw)) + C::kUsedBytes > (uint32_t)C::kSmemBytes) __trap();       // alignment slack exhausted
    uint64_t* empty_bar = full_bar + C::kStages;
    uint64_t* tfull_bar = empty_bar + C::kStages;      // [2] accumulator ready
    uint64_t* tempty_bar = tfull_bar + 2;              // [2] accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int cta_rank = (CG == 2) ? (int)cluster_ctarank() : 0;
    const int first_tile = blockIdx.x / CG;            // tiles are dealt to CTAs (CG = 1) or CTA pairs (CG = 2)
    const int tile_step = gridDim.x / CG;
    const int num_m = (p.M + BM * CG - 1) / (BM * CG);
    const int num_n = p.N / BN;
    const int num_tiles = num_m * num_n;
    const int num_kb = p.K / BK;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], kEpiWarps * CG);   // CG = 2: the epilogue warps of both CTAs report to rank 0
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        if (CG == 2) tmem_alloc_pair<C::kTmemCols>(tmem_slot);
        else tmem_alloc<C::kTmemCols>(tmem_slot);
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all();                     // the peer's barriers exist before anything signals them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (the whole warp, convergent: ptx.cuh)
        // A K block must be refilled within (stages - 1) MMA groups of its slot being freed, and this warp shares its
        // scheduler with two epilogue warps: the loop below is ~20 instructions per K block, no division, no MUFU.
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
            const int m0 = (tile / num_n) * (BM * CG) + cta_rank * BM;
            const int n0 = (tile % num_n) * BN;
            int acol = 0, arow = m0;                       // K index kk reads A at (row + kk / a_cols, kk % a_cols)
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * C::kStageBytes;
                uint8_t* sb = sa + C::kABytes;
                const int kk = kb * BK;
                if (CG == 2) {
                    // both CTAs' bytes are counted on rank 0's barrier, which rank 0 arms for the pair
                    if (cta_rank == 0) mbar_expect_tx_elect(&full_bar[stage], 2 * C::kStageBytes);
                    tma_load_2d_pair_elect(sa, &tmap_a, &full_bar[stage], acol, arow);
                    tma_load_2d_pair_elect(sb, &tmap_b, &full_bar[stage], kk, n0 + cta_rank * kBBoxRows);
                } else {
                    mbar_expect_tx_elect(&full_bar[stage], C::kStageBytes);
                    tma_load_2d_elect(sa, &tmap_a, &full_bar[stage], acol, arow);
#pragma unroll
                    for (int h = 0; h < BN / kBBoxRows; ++h)
                        tma_load_2d_elect(sb + h * kBBoxRows * BK * 2, &tmap_b, &full_bar[stage], kk, n0 + h * kBBoxRows);
                }
                acol += BK;
                if (acol == p.a_cols) {
                    acol = 0;
                    ++arow;
                }
                if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (cta_rank == 0) {
            // ------------------------------------------------------------ MMA issuer (rank 0 of a pair issues for both)
            // The whole warp runs this loop convergently and every tcgen05 instruction is predicated on elect.sync
            // (ptx.cuh): under a divergent `if (lane == 0)` ptxas wraps each one in a loop over the active lanes,
            // ~13 instructions and ~80 cycles of issue per MMA -- more than half of the 135 cycles a 128x256x16 MMA
            // occupies the tensor pipe.  The four K = 16 steps of a stage go out as one statement.
            static_assert(BK == 64, "one x4 group per stage");
            constexpr bool kLnFold = (EPI == EPI_LN_GELU_BF16 || EPI == EPI_LN_QKV_SPLIT_BF16);
            constexpr uint32_t idesc = kLnFold ? umma_idesc_f16(BM * CG, BN) : umma_idesc_bf16(BM * CG, BN, false, false);
            constexpr uint64_t desc_hi64 = umma_smem_desc_hi(16, 1024);   // K-major SW128: SBO = 8 rows * 128 B
            constexpr uint32_t desc_hi = (uint32_t)(desc_hi64 >> 32);
            const uint32_t desc_lo0 = (uint32_t)(desc_hi64 & 0xFFFFFFFFu) | ((base >> 4) & 0x3FFF);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                const bool tr = p.trace != nullptr && blockIdx.x == 0 && it < p.trace_tiles;
                unsigned long long t0 = 0, t_wait = 0;
                if (tr) t0 = clock64();
                mbar_wait(&tempty_bar[as], aphase ^ 1);
                tc_fence_after();
                if (tr && lane == 0) { p.trace[it * 8 + 0] = t0; p.trace[it * 8 + 1] = clock64(); }
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    unsigned long long w0 = 0;
                    if (tr) w0 = clock64();
                    mbar_wait(&full_bar[stage], phase);
                    if (tr) t_wait += clock64() - w0;
                    tc_fence_after();
                    const uint32_t a_lo = desc_lo0 + ((uint32_t)(stage * C::kStageBytes) >> 4);
                    if (CG == 2) {
                        umma_bf16_ss_x4_elect_pair(d_tmem, a_lo, a_lo + (C::kABytes >> 4), desc_hi, idesc, kb != 0);
                        umma_commit_elect_pair(&empty_bar[stage]);           // frees the slot in BOTH CTAs
                        if (kb == num_kb - 1) umma_commit_elect_pair(&tfull_bar[as]);
                    } else {
                        umma_bf16_ss_x4_elect(d_tmem, a_lo, a_lo + (C::kABytes >> 4), desc_hi, idesc, kb != 0);
                        umma_commit_elect(&empty_bar[stage]);    // frees the smem slot once these MMAs retire
                        if (kb == num_kb - 1) umma_commit_elect(&tfull_bar[as]);
                    }
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
                if (tr && lane == 0) { p.trace[it * 8 + 2] = t_wait; p.trace[it * 8 + 3] = clock64(); }
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue warps
        // TMEM hands every thread one accumulator ROW (32 consecutive f32 columns per load).  Every output of this
        // kernel is a 2-byte type, so those 32 columns are 64 contiguous bytes = two full 32-byte sectors of the thread's
        // own row: each thread stores them with two 256-bit STG (and reads the f16 residual the same way), a warp-level
        // access = 32 full sectors.  Round 1 transposed every 32x32 chunk through shared memory instead (needed when the
        // residual stream was f32); the round-2 timeline (tests/gemm_trace.py) showed what that cost: the staging
        // traffic shares the shared-memory port with the MMA's operand reads, and a K = 64 block took 632 cycles in fc1
        // (epilogue active 54 % of the time) against 540 in fc2 (24 %) for the same 512-cycle MMA work.
        const int ew = warp - 2;
        const int quarter = warp & 3;          // TMEM lane quarter this warp may touch
        const int half = ew >> 2;              // which half of the tile's columns
        constexpr int kChunks = (BN / 2) / 32;
        constexpr bool kLn = (EPI == EPI_LN_GELU_BF16 || EPI == EPI_LN_QKV_SPLIT_BF16);
        constexpr bool kQkv = (EPI == EPI_QKV_SPLIT_BF16 || EPI == EPI_LN_QKV_SPLIT_BF16);
        constexpr bool kGelu = (EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_GELU_POS_F16 || EPI == EPI_LN_GELU_BF16);
        constexpr bool kResid = (EPI == EPI_BIAS_RESID_F16);
        constexpr bool kPos = (EPI == EPI_BIAS_GELU_POS_F16);
        constexpr bool kF16Out = kResid || kPos;
        // Per-tile inputs that do not depend on the accumulator are fetched ONE TILE AHEAD, so that their L2 latency hides
        // under the previous tile's drain instead of opening every tile (round-2 timeline: ~2700 of fc1's ~11300 cycles
        // per tile): the row's LayerNorm partials (registers -> -rstd mean, rstd) and the tile's bias / c1 values (one
        // register each, parked in shared memory at the next tile's start).
        const int e_idx = threadIdx.x - 64;                              // 0 .. 255 among the epilogue threads
        auto row_of = [&](int tile) { return (tile / num_n) * (BM * CG) + cta_rank * BM + quarter * 32 + lane; };
        constexpr int kMaxParts = 16;
        float2 pvn[kLn ? kMaxParts : 1];
        auto stats_issue = [&](int tile) {                                   // loads only; consumed by stats_finish
            const int r = row_of(tile);
#pragma unroll
            for (int k = 0; k < kMaxParts; ++k)
                if (kLn) pvn[k] = (k < p.stats_parts && r < p.M) ? __ldg(p.stats_in + (long long)r * p.stats_parts + k)
                                                                 : make_float2(0.f, 0.f);
        };
        auto stats_finish = [&](float& nmr_out, float& rstd_out) {
            float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
            for (int k = 0; k < kMaxParts; ++k) {
                if (kLn) {
                    s1 += pvn[k].x;
                    s2 += pvn[k].y;
                }
            }
            const float inv_d = 1.0f / (float)p.ln_dim;
            const float mean = s1 * inv_d;
            rstd_out = rsqrtf(fmaxf(s2 * inv_d - mean * mean, 0.0f) + p.ln_eps);
            nmr_out = -mean * rstd_out;
        };
        float nmr = 0.0f, rstd = 1.0f;
        float vb_next = 0.0f, vc_next = 0.0f;
        if (first_tile < num_tiles) {
            if (kLn) {
                stats_issue(first_tile);
                stats_finish(nmr, rstd);
            }
            if (e_idx < BN) {
                vb_next = __ldg(p.bias + (first_tile % num_n) * BN + e_idx);
                if (kLn) vc_next = __ldg(p.c1 + (first_tile % num_n) * BN + e_idx);
            }
        }
        int it = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int m0 = (tile / num_n) * (BM * CG) + cta_rank * BM;
            const int n0 = (tile % num_n) * BN + half * (BN / 2);
            const int next_tile = tile + tile_step;
            const bool has_next = next_tile < num_tiles;
            // this thread's accumulator row -> output row
            const int r_own = m0 + quarter * 32 + lane;
            const int b_own = r_own / p.p_in;
            const int t_own = r_own - b_own * p.p_in;
            const bool valid_own = (r_own < p.M) && (t_own < p.t_valid);
            const long long orow = valid_own ? ((long long)b_own * p.p_out + t_own + p.row_off) : 0;
            const long long obase = orow * p.ldo;                 // element offset of the row start (rows never written: row 0, loads stay in bounds)

            // LayerNorm fold, consuming side: (nmr, rstd) = (-rstd mean, rstd) of the own row were prepared one tile ahead
            // LayerNorm fold, producing side: (sum, sum of squares) of what this thread stores of its row
            const bool produce = (kResid || kPos) && p.stats_out != nullptr;
            float st_s = 0.0f, st_q = 0.0f;

            // the residual / position values do not depend on the accumulator: chunk 0's are fetched before waiting for
            // the MMA and chunk c + 1's while chunk c is processed, so their DRAM latency is off the critical path
            // (f16 residual: all of the tile's chunks are requested up front -- 64 registers -- so that the chunk loop holds
            //  no global loads at all: with one chunk of lookahead the tcgen05.ld of the next chunk queued behind the
            //  outstanding residual loads on the scoreboard, 19 % of the out-projection's samples in round 2)
            constexpr int kAddBufs = kResid ? kChunks : 2;
            uint32_t addr[kAddBufs][kPos ? 32 : 16];
            auto load_add = [&](int c, uint32_t (&dst)[kPos ? 32 : 16]) {
                const int nc = n0 + c * 32;
                if (kResid) {
                    const __half* src = reinterpret_cast<const __half*>(p.resid) + obase + nc;
                    ldg256(src, dst[0], dst[1], dst[2], dst[3], dst[4], dst[5], dst[6], dst[7]);
                    ldg256(src + 16, dst[8], dst[9], dst[10], dst[11], dst[12], dst[13], dst[14], dst[15]);
                } else if (kPos) {
                    const float* src = p.pos + (long long)(valid_own ? t_own : 0) * p.N + nc;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        ldg256_nc(src + 8 * q, dst[8 * q + 0], dst[8 * q + 1], dst[8 * q + 2], dst[8 * q + 3], dst[8 * q + 4],
                                  dst[8 * q + 5], dst[8 * q + 6], dst[8 * q + 7]);
                }
            };
            if (kResid) {
#pragma unroll
                for (int c = 0; c < kChunks; ++c) load_add(c, addr[c]);
            } else if (kPos) {
                load_add(0, addr[0]);
            }

            // the tile's per-column vectors go through shared memory: 256 epilogue threads hold one bias (and c1) value each,
            // fetched coalesced during the PREVIOUS tile; the chunk loop reads them as broadcast LDS.128.  (Fetched with
            // warp-uniform __ldg per chunk they missed the ~25 KB of L1 left beside 227 KB of shared memory and their L2
            // latency sat on every chunk's critical path: 16 % of fc1's samples in round 2.)
            asm volatile("bar.sync 1, 256;" ::: "memory");                   // every warp is done with the previous tile's vectors
            if (e_idx < BN) {
                vec[e_idx] = vb_next;
                if (kLn) vec[BN + e_idx] = vc_next;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (has_next && e_idx < BN) {
                vb_next = __ldg(p.bias + (next_tile % num_n) * BN + e_idx);
                if (kLn) vc_next = __ldg(p.c1 + (next_tile % num_n) * BN + e_idx);
            }
            const float4* vbias = reinterpret_cast<const float4*>(vec + half * (BN / 2));
            const float4* vc1 = reinterpret_cast<const float4*>(vec + BN + half * (BN / 2));

            const bool etr = p.trace != nullptr && blockIdx.x == 0 && ew == 0 && lane == 0 && it < p.trace_tiles;
            if (etr) p.trace[it * 8 + 4] = clock64();
            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after();
            if (etr) p.trace[it * 8 + 5] = clock64();
            if (kLn && has_next) stats_issue(next_tile);
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                uint32_t acc[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * BN + half * (BN / 2) + c * 32;
                tmem_ld_32x32b_x32(taddr, acc);
                const int nc = n0 + c * 32;
                const float4* bias4 = vbias + c * 8;                     // shared memory, broadcast reads
                const float4* c1_4 = vc1 + c * 8;
                // the chunk's bias (and c1) values are read under the tcgen05.ld's latency: left to ptxas the LDS sit right
                // in front of their first use (short-scoreboard stalls: 24 % of fc1's samples in round 2); 320 threads per
                // SM leave 204 registers per thread, so the 64 extra registers cost no occupancy
                float4 bbv[8], ccv[kLn ? 8 : 1];
                if (!(kQkv && nc >= p.n_split)) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        bbv[j] = bias4[j];
                        if (kLn) ccv[j] = c1_4[j];
                    }
                }
                tmem_ld_wait_on(acc);
                if (kQkv && nc >= p.n_split) {
                    // values: out2[b][head][c][t]; for a fixed column the warp's 32 rows are 32 consecutive t
                    if (valid_own) {
                        __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(p.out2) +
                                            ((long long)b_own * (p.N - p.n_split) + (nc - p.n_split)) * p.t_pad + t_own;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 bb = bias4[j];
                            float4 a4 = make_float4(__uint_as_float(acc[4 * j + 0]), __uint_as_float(acc[4 * j + 1]),
                                                    __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]));
                            if (kLn) {
                                const float4 cc = c1_4[j];
                                a4.x = fmaf(rstd, a4.x, fmaf(nmr, cc.x, bb.x));
                                a4.y = fmaf(rstd, a4.y, fmaf(nmr, cc.y, bb.y));
                                a4.z = fmaf(rstd, a4.z, fmaf(nmr, cc.z, bb.z));
                                a4.w = fmaf(rstd, a4.w, fmaf(nmr, cc.w, bb.w));
                            } else {
                                a4.x += bb.x; a4.y += bb.y; a4.z += bb.z; a4.w += bb.w;
                            }
                            o2[(long long)(4 * j + 0) * p.t_pad] = __float2bfloat16_rn(a4.x);
                            o2[(long long)(4 * j + 1) * p.t_pad] = __float2bfloat16_rn(a4.y);
                            o2[(long long)(4 * j + 2) * p.t_pad] = __float2bfloat16_rn(a4.z);
                            o2[(long long)(4 * j + 3) * p.t_pad] = __float2bfloat16_rn(a4.w);
                        }
                    }
                    continue;
                }
                if (kPos && c + 1 < kChunks) load_add(c + 1, addr[(c + 1) & 1]);
                uint32_t pk[16];
                const float2 rs2 = make_float2(rstd, rstd), nm2 = make_float2(nmr, nmr);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 bb = bbv[j];
                    float2 lo = make_float2(__uint_as_float(acc[4 * j + 0]), __uint_as_float(acc[4 * j + 1]));
                    float2 hi = make_float2(__uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]));
                    if (kLn) {
                        // rstd acc + (c2 - rstd mean c1): two packed FMAs per pair, bias (= c2, in `bb`) included
                        const float4 cc = ccv[kLn ? j : 0];
                        lo = __ffma2_rn(rs2, lo, __ffma2_rn(nm2, make_float2(cc.x, cc.y), make_float2(bb.x, bb.y)));
                        hi = __ffma2_rn(rs2, hi, __ffma2_rn(nm2, make_float2(cc.z, cc.w), make_float2(bb.z, bb.w)));
                    } else {
                        lo = __fadd2_rn(lo, make_float2(bb.x, bb.y));
                        hi = __fadd2_rn(hi, make_float2(bb.z, bb.w));
                    }
                    if (kGelu) {
                        lo = gelu_poly2(lo);
                        hi = gelu_poly2(hi);
                    }
                    if (!kF16Out) {
                        pk[2 * j] = pack_bf16x2(lo.x, lo.y);
                        pk[2 * j + 1] = pack_bf16x2(hi.x, hi.y);
                    } else {
                        float a0, a1, a2, a3;
                        if (kResid) {
                            const float2 x0 = __half22float2(*reinterpret_cast<const __half2*>(&addr[c][2 * j]));
                            const float2 x1 = __half22float2(*reinterpret_cast<const __half2*>(&addr[c][2 * j + 1]));
                            a0 = x0.x; a1 = x0.y; a2 = x1.x; a3 = x1.y;
                        } else {
                            a0 = __uint_as_float(addr[c & 1][4 * j + 0]); a1 = __uint_as_float(addr[c & 1][4 * j + 1]);
                            a2 = __uint_as_float(addr[c & 1][4 * j + 2]); a3 = __uint_as_float(addr[c & 1][4 * j + 3]);
                        }
                        // residual stream in f16 (as CTranslate2's float16 mode keeps it): saturate instead of inf
                        const float kMaxHalf = 65504.0f;
                        const float r0 = fminf(fmaxf(lo.x + a0, -kMaxHalf), kMaxHalf);
                        const float r1 = fminf(fmaxf(lo.y + a1, -kMaxHalf), kMaxHalf);
                        const float r2 = fminf(fmaxf(hi.x + a2, -kMaxHalf), kMaxHalf);
                        const float r3 = fminf(fmaxf(hi.y + a3, -kMaxHalf), kMaxHalf);
                        const __half2 h0 = __floats2half2_rn(r0, r1), h1 = __floats2half2_rn(r2, r3);
                        pk[2 * j] = *reinterpret_cast<const unsigned*>(&h0);
                        pk[2 * j + 1] = *reinterpret_cast<const unsigned*>(&h1);
                        st_s += (r0 + r1) + (r2 + r3);
                        st_q = fmaf(r0, r0, fmaf(r1, r1, fmaf(r2, r2, fmaf(r3, r3, st_q))));
                    }
                }
                if (valid_own) {
                    uint16_t* dst = reinterpret_cast<uint16_t*>(p.out) + obase + nc;
                    stg256(dst, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
                    stg256(dst + 16, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
                }
            }
            if (kLn && has_next) stats_finish(nmr, rstd);                   // for the next tile
            if (produce && valid_own) {
                // this thread owns BN / 2 columns of its row: one (sum, sum of squares) slice, no cross-lane reduction
                const int parts = p.N / (BN / 2);
                const int slice = (tile % num_n) * 2 + half;
                p.stats_out[orow * parts + slice] = make_float2(st_s, st_q);
            }
            tc_fence_before();
            __syncwarp();
            if (etr) p.trace[it * 8 + 6] = clock64();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_rank0(&tempty_bar[as]);
                else mbar_arrive(&tempty_bar[as]);
            }
        }
    }

    tc_fence_before();
    if (CG == 2) cluster_sync_all();                     // neither CTA may retire while its peer still uses it
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_pair<C::kTmemCols>(tmem_base);
        else tmem_dealloc<C::kTmemCols>(tmem_base);
    }
}

bool use_cta_pairs() {                     // ARIES_GEMM_PAIR=0 falls back to the 1-CTA kernel (A/B comparisons)
    static const bool on = [] {
        const char* e = getenv("ARIES_GEMM_PAIR");
        return e ? (e[0] != '0') : kDefaultPair;
    }();
    return on;
}

template <int BN, int EPI>
cudaError_t launch_one(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int sm_count,
                       cudaStream_t stream) {
    if (BN == 256 && use_cta_pairs() && p.M > BM) {
        using C = Cfg<256, 2>;
        const int tiles = ((p.M + 2 * BM - 1) / (2 * BM)) * (p.N / 256);
        const int pairs = tiles < sm_count / 2 ? tiles : sm_count / 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = C::kSmemBytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05<256, EPI, 2>, ta, tb, p);
    }
    using C = Cfg<BN, 1>;
    auto kern = gemm_bf16_tcgen05<BN, EPI, 1>;
    const int tiles = ((p.M + BM - 1) / BM) * (p.N / BN);
    const int grid = tiles < sm_count ? tiles : sm_count;
    kern<<<grid, kThreads, C::kSmemBytes, stream>>>(ta, tb, p);
    return cudaGetLastError();
}

template <int BN>
cudaError_t launch_bn(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int sm_count,
                      cudaStream_t stream) {
    switch (epi) {
        case EPI_BIAS_BF16: return launch_one<BN, EPI_BIAS_BF16>(ta, tb, p, sm_count, stream);
        case EPI_BIAS_GELU_BF16: return launch_one<BN, EPI_BIAS_GELU_BF16>(ta, tb, p, sm_count, stream);
        case EPI_BIAS_RESID_F16: return launch_one<BN, EPI_BIAS_RESID_F16>(ta, tb, p, sm_count, stream);
        case EPI_BIAS_GELU_POS_F16: return launch_one<BN, EPI_BIAS_GELU_POS_F16>(ta, tb, p, sm_count, stream);
        case EPI_QKV_SPLIT_BF16: return launch_one<BN, EPI_QKV_SPLIT_BF16>(ta, tb, p, sm_count, stream);
        case EPI_LN_GELU_BF16: return launch_one<BN, EPI_LN_GELU_BF16>(ta, tb, p, sm_count, stream);
        case EPI_LN_QKV_SPLIT_BF16: return launch_one<BN, EPI_LN_QKV_SPLIT_BF16>(ta, tb, p, sm_count, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace

int gemm_block_n(int N) { return (N % 256 == 0) ? 256 : 128; }
int gemm_b_box_rows() { return kBBoxRows; }

namespace {
template <int BN, int EPI>
cudaError_t set_smem() {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05<BN, EPI, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg<BN, 1>::kSmemBytes);
    if (e != cudaSuccess || BN != 256) return e;
    return cudaFuncSetAttribute(gemm_bf16_tcgen05<256, EPI, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                Cfg<256, 2>::kSmemBytes);
}
template <int BN>
cudaError_t set_smem_bn() {
    cudaError_t e;
    if ((e = set_smem<BN, EPI_BIAS_BF16>()) != cudaSuccess) return e;
    if ((e = set_smem<BN, EPI_BIAS_GELU_BF16>()) != cudaSuccess) return e;
    if ((e = set_smem<BN, EPI_BIAS_RESID_F16>()) != cudaSuccess) return e;
    if ((e = set_smem<BN, EPI_QKV_SPLIT_BF16>()) != cudaSuccess) return e;
    if ((e = set_smem<BN, EPI_LN_GELU_BF16>()) != cudaSuccess) return e;
    if ((e = set_smem<BN, EPI_LN_QKV_SPLIT_BF16>()) != cudaSuccess) return e;
    return set_smem<BN, EPI_BIAS_GELU_POS_F16>();
}
}  // namespace

// Per device (call after cudaSetDevice): opt in to > 48 KB dynamic shared memory for every instance.
cudaError_t gemm_init_device() {
    cudaError_t e = set_smem_bn<256>();
    if (e != cudaSuccess) return e;
    return set_smem_bn<128>();
}

cudaError_t gemm_launch(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int sm_count,
                        cudaStream_t stream) {
    if (p.K % BK != 0 || p.N % 128 != 0 || p.a_cols % BK != 0 || p.M <= 0) return cudaErrorInvalidValue;
    if ((epi == EPI_LN_GELU_BF16 || epi == EPI_LN_QKV_SPLIT_BF16) &&
        (!p.stats_in || !p.c1 || p.stats_parts <= 0 || p.stats_parts > 16 || p.ln_dim <= 0))
        return cudaErrorInvalidValue;
    return gemm_block_n(p.N) == 256 ? launch_bn<256>(epi, ta, tb, p, sm_count, stream)
                                    : launch_bn<128>(epi, ta, tb, p, sm_count, stream);
}

}  // namespace aries
