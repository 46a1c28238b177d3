// pack.cuh -- K9: conversions between the reference's host byte layouts and the
// packed device layouts (upload, download, CSV export).
//   core:      one-hot byte {1,2,4,8} (population.rs:201-204)  <->  2-bit code {0,1,2,3}
//   accessory: byte {0,1} (population.rs:214-219)              <->  1 bit
#pragma once
#include "common.cuh"

namespace pansim {

// one-hot -> code: 1->0, 2->1, 4->2, 8->3
__device__ __forceinline__ uint32_t onehot_to_code(uint32_t b) { return (b >> 1) - (b >> 3); }

// bytes [n_rows x n_sites] (row chunk staged on the device) -> packed rows.
// One thread per packed 32-bit word (16 sites). Sets *bad if a byte is not one-hot.
__global__ void pack_core_kernel(const uint8_t *bytes, uint64_t n_sites, uint32_t row0, uint32_t n_rows,
                                 uint8_t *state, uint64_t row_stride, int *bad)
{
    const uint64_t words_per_row = row_stride / 4;
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)n_rows * words_per_row) return;
    const uint32_t r = (uint32_t)(idx / words_per_row);
    const uint64_t w = idx % words_per_row;
    const uint8_t *src = bytes + (uint64_t)r * n_sites;
    uint32_t out = 0;
    for (uint32_t k = 0; k < 16; k++) {
        const uint64_t s = w * 16 + k;
        if (s < n_sites) {
            const uint32_t b = src[s];
            if (!(b == 1 || b == 2 || b == 4 || b == 8)) *bad = 1;
            out |= (onehot_to_code(b) & 3u) << (2 * k);
        }
    }
    reinterpret_cast<uint32_t *>(state + (uint64_t)(row0 + r) * row_stride)[w] = out;
}

__global__ void unpack_core_kernel(const uint8_t *state, uint64_t row_stride, uint64_t n_sites, uint32_t row0,
                                   uint32_t n_rows, uint8_t *bytes)
{
    const uint64_t words_per_row = (n_sites + 15) / 16;
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)n_rows * words_per_row) return;
    const uint32_t r = (uint32_t)(idx / words_per_row);
    const uint64_t w = idx % words_per_row;
    const uint32_t v = reinterpret_cast<const uint32_t *>(state + (uint64_t)(row0 + r) * row_stride)[w];
    uint8_t *dst = bytes + (uint64_t)r * n_sites;
    for (uint32_t k = 0; k < 16; k++) {
        const uint64_t s = w * 16 + k;
        if (s < n_sites) dst[s] = (uint8_t)(1u << ((v >> (2 * k)) & 3u));
    }
}

// `A,C,G,T\n` rows of _core_genome.csv (population.rs:877-879, int_to_base :154-162)
__global__ void export_core_csv_kernel(const uint8_t *state, uint64_t row_stride, uint64_t n_sites,
                                       uint32_t row0, uint32_t n_rows, char *out)
{
    const uint64_t words_per_row = (n_sites + 15) / 16;
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)n_rows * words_per_row) return;
    const uint32_t r = (uint32_t)(idx / words_per_row);
    const uint64_t w = idx % words_per_row;
    const uint32_t v = reinterpret_cast<const uint32_t *>(state + (uint64_t)(row0 + r) * row_stride)[w];
    char *dst = out + (uint64_t)r * (2 * n_sites);
    const uint32_t lut = ('A') | ('C' << 8) | ('G' << 16) | ('T' << 24);
    for (uint32_t k = 0; k < 16; k++) {
        const uint64_t s = w * 16 + k;
        if (s < n_sites) {
            dst[2 * s] = (char)((lut >> (8 * ((v >> (2 * k)) & 3u))) & 0xFF);
            dst[2 * s + 1] = (s + 1 == n_sites) ? '\n' : ',';
        }
    }
}

// copy packed row 0 into rows 1..n_rows-1 (population.rs:206-212: clonal start)
__global__ void replicate_row_kernel(uint8_t *state, uint64_t row_stride, uint32_t n_rows)
{
    const uint64_t vec_per_row = row_stride / 16;
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)(n_rows - 1) * vec_per_row) return;
    const uint64_t r = 1 + idx / vec_per_row, v = idx % vec_per_row;
    reinterpret_cast<uint4 *>(state + r * row_stride)[v] = reinterpret_cast<const uint4 *>(state)[v];
}

__global__ void pack_acc_kernel(const uint8_t *bytes, uint32_t n_genes, uint32_t n_rows, uint32_t *state,
                                uint32_t stride_words, int *bad)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)n_rows * stride_words) return;
    const uint32_t r = (uint32_t)(idx / stride_words), w = (uint32_t)(idx % stride_words);
    uint32_t out = 0;
    for (uint32_t k = 0; k < 32; k++) {
        const uint32_t g = w * 32 + k;
        if (g < n_genes) {
            const uint32_t b = bytes[(uint64_t)r * n_genes + g];
            if (b > 1) *bad = 1;
            out |= (b & 1u) << k;
        }
    }
    state[idx] = out;
}

__global__ void unpack_acc_kernel(const uint32_t *state, uint32_t stride_words, uint32_t n_genes,
                                  uint32_t n_rows, uint8_t *bytes)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)n_rows * n_genes) return;
    const uint32_t r = (uint32_t)(idx / n_genes), g = (uint32_t)(idx % n_genes);
    bytes[idx] = (uint8_t)((state[(uint64_t)r * stride_words + (g >> 5)] >> (g & 31u)) & 1u);
}

}  // namespace pansim
