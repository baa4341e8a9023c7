// Tensor-core attention for LONG key sets with head_dim 32 (the co-evolution decoder's vertex self-attention, 431 x 431,
// and its vertex<-joint cross-attention, 431 x 17): flash-style loop over 128-key chunks with the running max / sum and
// the output row kept in registers.
//   item  = (sequence b, head h, 128-query tile); per 128-key chunk:
//   S  = (Q*scale) K^T : tcgen05 128 x 128, split-bf16. head_dim 32 gives 64-byte operand rows, so hi|lo are CONCATENATED
//        into one 128-byte swizzled row: A = [Q_hi|Q_lo], B1 = [K_hi|K_hi] (k16 steps 0-3: Q_hi K_hi + Q_lo K_hi),
//        B2 = [K_lo| - ] (steps 0-1: Q_hi K_lo)  -> the same 6 MMAs a 3-term split needs, one swizzle mode everywhere.
//   P  = exp(S - m_new) (keys beyond N2 masked), split-bf16 -> smem as two [128][64] K-major tiles (hi) + two (lo),
//        overwriting the chunk's K tiles (+ two spare tiles) once S is complete.
//   D  = P_hi [V_hi|V_lo] + P_lo [V_hi|V_lo] : tcgen05 128 x 64 x 128 with V consumed as an MN-major B operand;
//        columns 0-31 + columns 32-63 = this chunk's P V (the extra P_lo V_lo term only adds accuracy).
//   o  = o * exp(m_old - m_new) + D ; after the last chunk o / l -> split-bf16 -> global.
// Same skeleton as attn_tc.cuh: two independent math groups (warps 0-3 / 4-7; own smem, TMEM, barriers) alternate items,
// warps 8-15 load fp32 q/k/v rows, split them and write the swizzled tiles.
#pragma once
#include "tc_common.cuh"
#include "common.cuh"
#include "kernels.cuh"
#include "attn_tc.cuh"

constexpr int AF_BUF = 6 * AT_TILE;                // Qcat | Khh | Kl | X1 | X2 | Vcat   (P hi = Khh|Kl, P lo = X1|X2)
constexpr int AF_SMEM = 2 * AF_BUF + 1024 + 128;
constexpr int AF_THREADS = 512;

__global__ void __launch_bounds__(AF_THREADS, 1)
attn_flash_tc_kernel(const float* __restrict__ Q, AttnAddr aq, const float* __restrict__ K, const float* __restrict__ V, AttnAddr akv, SplitOut Os,
                     AttnAddr ao, int N1, int N2, int nseq, int H, float scale) {
    constexpr int D = 32;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sb = tc::smem_u32(smem);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + 2 * AF_BUF);   // [2] per group, 256 loader arrivals per chunk
    uint64_t* empty_bar = full_bar + 2;                                    // [2] per group, 1 arrival (tcgen05.commit of the PV MMAs)
    uint64_t* bar_s = empty_bar + 2;
    uint64_t* bar_o = bar_s + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_o + 2);

    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&full_bar[b], 256); tc::mbar_init(&empty_bar[b], 1);
            tc::mbar_init(&bar_s[b], 1); tc::mbar_init(&bar_o[b], 1);
        }
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_ptr_smem, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int qtiles = (N1 + 127) / 128, nchunks = (N2 + 127) / 128;
    const int nwork = nseq * H * qtiles;           // work -> (seq, head, qtile), qtile fastest (K/V of a (seq, head) stay L2-hot)

    if (warp >= 8) {
        // ================= loaders: thread = (row, half of the 32-wide head slice) =================
        const int lt = tid - 256;
        const int lr = lt & 127, hf = lt >> 7;
        uint32_t n = 0;       // item counter of this CTA
        uint32_t fill0 = 0, fill1 = 0;   // per-group chunk-fill counters (mbarrier phases)
        for (int work = blockIdx.x; work < nwork; work += gridDim.x, ++n) {
            const int grp = n & 1;
            const uint32_t base_s = sb + grp * AF_BUF;
            const uint32_t Qc = base_s, Khh = base_s + AT_TILE, Kl = base_s + 2 * AT_TILE, Vc = base_s + 5 * AT_TILE;
            const int qt = work % qtiles, sh = work / qtiles;
            const int h = sh % H, s = sh / H;
            for (int c = 0; c < nchunks; ++c) {
                const int key = c * 128 + lr;
                const bool kvalid = key < N2;
                float4 kx[4], vx[4], qx[4];
                if (kvalid) {
                    const size_t kb = (size_t)(akv.seq(s) + (long long)key * akv.tok) * akv.ld + h * D + hf * 16;
#pragma unroll
                    for (int i = 0; i < 4; ++i) { kx[i] = ld4(K + kb + i * 4); vx[i] = ld4(V + kb + i * 4); }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) kx[i] = vx[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                const int qrow = qt * 128 + lr;
                if (c == 0) {
                    if (qrow < N1) {
                        const size_t qb = (size_t)(aq.seq(s) + (long long)qrow * aq.tok) * aq.ld + h * D + hf * 16;
#pragma unroll
                        for (int i = 0; i < 4; ++i) qx[i] = ld4(Q + qb + i * 4);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) qx[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                const uint32_t fl = grp ? fill1 : fill0;
                tc::mbar_wait(&empty_bar[grp], (fl & 1) ^ 1);
                if (grp) ++fill1; else ++fill0;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {          // two 8-element chunks of this thread's 16 columns
                    float x[8];
                    uint4 hh, ll;
                    x[0] = kx[2 * cc].x; x[1] = kx[2 * cc].y; x[2] = kx[2 * cc].z; x[3] = kx[2 * cc].w;
                    x[4] = kx[2 * cc + 1].x; x[5] = kx[2 * cc + 1].y; x[6] = kx[2 * cc + 1].z; x[7] = kx[2 * cc + 1].w;
                    tc::split8(x, hh, ll);
                    tc::sts16(Khh, lr, hf * 2 + cc, hh); tc::sts16(Khh, lr, 4 + hf * 2 + cc, hh);    // [K_hi | K_hi]
                    tc::sts16(Kl, lr, hf * 2 + cc, ll);                                                // [K_lo |  -  ]
                    x[0] = vx[2 * cc].x; x[1] = vx[2 * cc].y; x[2] = vx[2 * cc].z; x[3] = vx[2 * cc].w;
                    x[4] = vx[2 * cc + 1].x; x[5] = vx[2 * cc + 1].y; x[6] = vx[2 * cc + 1].z; x[7] = vx[2 * cc + 1].w;
                    tc::split8(x, hh, ll);
                    tc::sts16(Vc, lr, hf * 2 + cc, hh); tc::sts16(Vc, lr, 4 + hf * 2 + cc, ll);        // [V_hi | V_lo]
                    if (c == 0) {
                        x[0] = qx[2 * cc].x * scale; x[1] = qx[2 * cc].y * scale; x[2] = qx[2 * cc].z * scale; x[3] = qx[2 * cc].w * scale;
                        x[4] = qx[2 * cc + 1].x * scale; x[5] = qx[2 * cc + 1].y * scale; x[6] = qx[2 * cc + 1].z * scale; x[7] = qx[2 * cc + 1].w * scale;
                        tc::split8(x, hh, ll);
                        tc::sts16(Qc, lr, hf * 2 + cc, hh); tc::sts16(Qc, lr, 4 + hf * 2 + cc, ll);    // [Q_hi | Q_lo]
                    }
                }
                tc::fence_proxy_async();
                tc::mbar_arrive(&full_bar[grp]);
            }
        }
    } else {
        // ================= math: two groups of 4 warps =================
        const int grp = warp >> 2;
        const int r = tid & 127;                       // query row of the tile owned by this thread
        const int wq = warp & 3;                       // TMEM lane group
        const uint32_t tS = tmem_base + grp * 256, tD = tS + 128;
        const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;
        const uint32_t base_s = sb + grp * AF_BUF;
        const uint32_t Qc = base_s, Khh = base_s + AT_TILE, Kl = base_s + 2 * AT_TILE, X1 = base_s + 3 * AT_TILE, Vc = base_s + 5 * AT_TILE;
        const uint32_t Ph = Khh, Pl = X1;              // P hi = tiles {Khh, Kl}, P lo = tiles {X1, X2}
        const bool issuer = (tid & 127) == 0;
        uint32_t steps = 0;                            // chunk-step counter of this group (mbarrier phases)
        for (int work = blockIdx.x + grp * gridDim.x; work < nwork; work += 2 * gridDim.x) {
            const int qt = work % qtiles, sh = work / qtiles;
            const int h = sh % H, s = sh / H;
            const int qrow = qt * 128 + r;
            float m = -INFINITY, l = 0.f;
            float o[D];
#pragma unroll
            for (int i = 0; i < D; ++i) o[i] = 0.f;
            for (int c = 0; c < nchunks; ++c, ++steps) {
                const uint32_t ph = steps & 1;
                tc::mbar_wait(&full_bar[grp], ph);
                tc::tc_fence_before();
                tc::bar_sync_group(1 + grp);                               // group done with TMEM S/D of the previous chunk
                if (issuer) {
                    tc::tc_fence_after();
                    constexpr uint32_t idesc = tc::umma_idesc_bf16_f32(128, 128);
                    const uint64_t qc = tc::umma_desc_sw128(Qc), khh = tc::umma_desc_sw128(Khh), kl = tc::umma_desc_sw128(Kl);
                    // Q_hi K_lo (steps 0,1 of [K_lo|-]) first, then [Q_hi|Q_lo] [K_hi|K_hi]
                    tc::umma_bf16(tS, tc::umma_desc_advance_k(qc, 0), tc::umma_desc_advance_k(kl, 0), idesc, 0);
                    tc::umma_bf16(tS, tc::umma_desc_advance_k(qc, 1), tc::umma_desc_advance_k(kl, 1), idesc, 1);
#pragma unroll
                    for (int k = 0; k < 4; ++k) tc::umma_bf16(tS, tc::umma_desc_advance_k(qc, k), tc::umma_desc_advance_k(khh, k), idesc, 1);
                    tc::umma_commit(&bar_s[grp]);
                }
                tc::mbar_wait(&bar_s[grp], ph);
                tc::tc_fence_after();

                // ---- online softmax over this chunk's keys [c*128, min(N2, c*128+128)) ----
                const int nk = N2 - c * 128 < 128 ? N2 - c * 128 : 128;
                float mc = -INFINITY;
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    if (cc * 32 >= nk) break;                                  // uniform
                    uint32_t v[32];
                    tc::tmem_ld_32x32(tS + lane_sel + cc * 32, v);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (cc * 32 + i < nk) mc = fmaxf(mc, __uint_as_float(v[i]));
                }
                const float m_new = fmaxf(m, mc);
                const float alpha = __expf(m - m_new);                         // first chunk: exp(-inf) = 0
                m = m_new;
                float lsum = 0.f;
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    const int t = cc >> 1;
                    if (cc * 32 >= nk) {                                       // uniform: fully masked 32 keys
                        const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) { tc::sts16(Ph + t * AT_TILE, r, (cc & 1) * 4 + jj, z); tc::sts16(Pl + t * AT_TILE, r, (cc & 1) * 4 + jj, z); }
                        continue;
                    }
                    uint32_t v[32];
                    tc::tmem_ld_32x32(tS + lane_sel + cc * 32, v);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        float p[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            p[i] = (cc * 32 + jj * 8 + i < nk) ? __expf(__uint_as_float(v[jj * 8 + i]) - m_new) : 0.f;
                            lsum += p[i];
                        }
                        uint4 hh, ll;
                        tc::split8(p, hh, ll);
                        tc::sts16(Ph + t * AT_TILE, r, (cc & 1) * 4 + jj, hh);
                        tc::sts16(Pl + t * AT_TILE, r, (cc & 1) * 4 + jj, ll);
                    }
                }
                l = l * alpha + lsum;
                tc::fence_proxy_async();
                tc::tc_fence_before();
                tc::bar_sync_group(1 + grp);
                if (issuer) {
                    tc::tc_fence_after();
                    constexpr uint32_t idesc = tc::umma_idesc_bf16_f32_bmn(128, 64);
                    const uint64_t vc = tc::umma_desc_sw128(Vc);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint64_t pl = tc::umma_desc_advance_k(tc::umma_desc_sw128(Pl + (k >> 2) * AT_TILE), k & 3);
                        tc::umma_bf16(tD, pl, vc + (uint64_t)(k * 2048 >> 4), idesc, k != 0);
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint64_t ph2 = tc::umma_desc_advance_k(tc::umma_desc_sw128(Ph + (k >> 2) * AT_TILE), k & 3);
                        tc::umma_bf16(tD, ph2, vc + (uint64_t)(k * 2048 >> 4), idesc, 1);
                    }
                    tc::umma_commit(&bar_o[grp]);
                    tc::umma_commit(&empty_bar[grp]);                       // K/V/P tiles free once these MMAs have read them
                }
                tc::mbar_wait(&bar_o[grp], ph);
                tc::tc_fence_after();
                {
                    uint32_t d0[32];
                    tc::tmem_ld_32x32(tD + lane_sel, d0);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < D; ++i) o[i] = o[i] * alpha + __uint_as_float(d0[i]);
                    tc::tmem_ld_32x32(tD + lane_sel + 32, d0);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < D; ++i) o[i] += __uint_as_float(d0[i]);
                }
            }
            if (qrow < N1) {
                const float inv = 1.0f / l;
                const size_t ob = (size_t)(ao.seq(s) + (long long)qrow * ao.tok) * ao.ld + h * D;
#pragma unroll
                for (int i = 0; i < D; i += 4) store_split4(Os, ob + i, make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv));
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

// nseq sequences, H heads of 32; queries N1 (addr aq), keys/values N2 (addr akv); split-bf16 output (addr ao).
static inline int launch_attn_flash_tc(const float* Q, AttnAddr aq, const float* K, const float* V, AttnAddr akv, SplitOut Os, AttnAddr ao, int nseq,
                                       int H, int N1, int N2, cudaStream_t st) {
    if (!pmce_configure_smem<attn_flash_tc_kernel>(AF_SMEM)) return 2;
    const long long work = (long long)nseq * H * ((N1 + 127) / 128);
    const int sms = tc_num_sms();
    const long long want = (work + 1) / 2;
    const int grid = (int)(want < sms ? (want < 1 ? 1 : want) : sms);
    attn_flash_tc_kernel<<<grid, AF_THREADS, AF_SMEM, st>>>(Q, aq, K, V, akv, Os, ao, N1, N2, nseq, H, 1.0f / sqrtf(32.0f));
    return cudaGetLastError() == cudaSuccess ? 0 : 3;
}
