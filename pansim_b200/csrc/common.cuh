// common.cuh -- device-side building blocks shared by all kernels:
// Philox4x32-10, Poisson-by-inversion tables,
// and the sm_100a PTX wrappers (mbarrier, cp.async.bulk = TMA 1-D bulk copies).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pansim {

// ---------------------------------------------------------------------------
// geometry of the packed core alignment
// ---------------------------------------------------------------------------
// 2 bits per site (A=0,C=1,G=2,T=3  <->  reference one-hot 1,2,4,8), site s of a
// row lives in 32-bit word s/16 at bit 2*(s%16). Rows are padded to whole
// REGIONs. A region is the unit one warp processes per pipeline stage and the
// unit the RNG is keyed on: lane l of the warp owns words {l, l+32, ..., l+480}
// of the region (bank-conflict-free in shared memory), i.e. 256 sites.
constexpr uint32_t REGION_BYTES = 2048;
constexpr uint32_t REGION_WORDS = REGION_BYTES / 4;      // 512
constexpr uint32_t REGION_SITES = REGION_BYTES * 4;      // 8192
constexpr uint32_t BLOCK_SITES = 256;                    // sites per lane per region
constexpr uint32_t WORDS_PER_LANE = 16;

// RNG stream ids (counter word 3, high half)
constexpr uint32_t STREAM_CORE_MUT = 1;
constexpr uint32_t STREAM_CORE_HR = 2;
constexpr uint32_t STREAM_ACC_FLIP = 3;
constexpr uint32_t STREAM_ACC_HGT = 4;
constexpr uint32_t STREAM_PARENTS = 5;

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11). Counter-based: output is a pure
// function of (key, counter), so results do not depend on thread/GPU count.
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2,
                                                      uint32_t &c3, uint32_t k0, uint32_t k1)
{
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key)
{
    uint32_t c0 = ctr.x, c1 = ctr.y, c2 = ctr.z, c3 = ctr.w;
    uint32_t k0 = key.x, k1 = key.y;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// The ten round keys of a launch, precomputed on the host and passed inside the kernel argument
// struct: they become constant-bank operands of the XORs instead of 20 additions per call.
struct PhiloxKeys { uint32_t k0[10], k1[10]; };

__host__ __device__ inline PhiloxKeys philox_key_schedule(uint2 key)
{
    PhiloxKeys rk;
    for (int r = 0; r < 10; r++) {
        rk.k0[r] = key.x + (uint32_t)r * 0x9E3779B9u;
        rk.k1[r] = key.y + (uint32_t)r * 0xBB67AE85u;
    }
    return rk;
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, const PhiloxKeys &rk)
{
    uint32_t c0 = ctr.x, c1 = ctr.y, c2 = ctr.z, c3 = ctr.w;
#pragma unroll
    for (int r = 0; r < 10; r++) philox_round(c0, c1, c2, c3, rk.k0[r], rk.k1[r]);
    return make_uint4(c0, c1, c2, c3);
}

// Core-genome streams (SNP positions / allele digits / counts, recombination sites / donors) use
// Philox4x32-7: the same bijection with seven rounds, the smallest round count of the family that is
// Crush-resistant (Salmon et al., SC'11, table 2: Philox4x32-7 passes BigCrush; ten rounds is the
// paper's safety margin). The core kernel is instruction-issue bound and makes 3.5 calls per lane per
// 2 KiB region, so three rounds fewer are ~5 % of its instructions. The accessory flip / HGT streams
// (two calls per 32-gene word, beside the core kernel on the same SMs) use the same seven rounds; the
// parent draws keep ten. PANSIM_CORE_PHILOX_ROUNDS=10 at compile time restores the margin everywhere.
#ifndef PANSIM_CORE_PHILOX_ROUNDS
#define PANSIM_CORE_PHILOX_ROUNDS 7
#endif
constexpr int CORE_PHILOX_ROUNDS = PANSIM_CORE_PHILOX_ROUNDS;

__host__ __device__ __forceinline__ uint4 philox_core(uint4 ctr, uint2 key)
{
    uint32_t c0 = ctr.x, c1 = ctr.y, c2 = ctr.z, c3 = ctr.w;
    uint32_t k0 = key.x, k1 = key.y;
#pragma unroll
    for (int r = 0; r < CORE_PHILOX_ROUNDS; r++) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint4 philox_core(uint4 ctr, const PhiloxKeys &rk)
{
    uint32_t c0 = ctr.x, c1 = ctr.y, c2 = ctr.z, c3 = ctr.w;
#pragma unroll
    for (int r = 0; r < CORE_PHILOX_ROUNDS; r++) philox_round(c0, c1, c2, c3, rk.k0[r], rk.k1[r]);
    return make_uint4(c0, c1, c2, c3);
}

// Counter layout used everywhere:  x = column block id (site block / gene word),
// y = individual (row), z = generation, w = (stream << 16) | refill index.
__host__ __device__ __forceinline__ uint4 make_ctr(uint32_t block, uint32_t row, uint32_t gen,
                                                   uint32_t stream)
{
    return make_uint4(block, row, gen, stream << 16);
}

// Poisson counts are drawn by CDF inversion of one 32-bit uniform: k = #{j : T[j] <= u} with
// T[j] = round(CDF(j) * 2^32) (built on the host in f64), padded with 0xFFFFFFFF. A mean above the
// table range is drawn as a sum of n_sub independent Poisson(mean / n_sub) (exact by additivity).
// core_mut.cuh (SNPs per 256-site block, with a guide table) and core_hr.cuh (recombination
// events per region, lane-parallel compare) hold the two samplers.
constexpr uint32_t POISSON_TABLE_MAX = 1024;

// ---------------------------------------------------------------------------
// sm_100a PTX: mbarrier + 1-D bulk async copies (TMA engine, UBLKCP in SASS)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}

// the same on a 32-bit shared-space address
__device__ __forceinline__ void mbar_wait_s(uint32_t bar_s, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar_s), "r"(parity)
            : "memory");
    } while (!ok);
}

// global -> shared bulk copy, completion signalled on an mbarrier (bytes % 16 == 0)
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// shared -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }

template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

template <int N>
__device__ __forceinline__ void bulk_wait()
{
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// Programmatic dependent launch (sm_90+). A kernel that calls pdl_launch_dependents() lets the next
// kernel of the stream be scheduled while it is still running; that kernel (launched with the
// programmatic-stream-serialization attribute) blocks in pdl_wait() until this grid has completed
// and its memory is visible. Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// make generic-proxy shared-memory writes visible to the async proxy (TMA)
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

}  // namespace pansim
