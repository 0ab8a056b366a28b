// TEST INFRASTRUCTURE ONLY -- runs the per-lane arithmetic of the product's CRC kernel on the host.
//
// There is no GPU in the build container, so the device functions of csrc/crc_kernels.cuh (they are __host__ __device__)
// are compiled here for the CPU and driven by loops that stand in for the 32 lanes of a warp, row by row exactly as
// crc_chunks<> does.  tests/test_crc_cpu.py compares the result with zlib (oracle/crc_oracle.py).  Nothing in the product
// links or loads this file; the product has no CPU checksum path for device buffers.
#include <stdint.h>
#include <string.h>

#include "crc_kernels.cuh"

extern "C" int crc_lanes_chunk_value(const uint8_t* chunk, uint32_t rows, const uint32_t* z512, const uint32_t* z4,
                                     const uint32_t* z16, uint32_t* out) {
    if (!chunk || !rows || rows > nvrx::kCrcChunkRows || !out) return 1;
    uint32_t t[32][4];
    memset(t, 0, sizeof(t));
    for (uint32_t r = 0; r < rows; ++r)
        for (uint32_t lane = 0; lane < 32; ++lane) {
            uint32_t w[4];
            memcpy(w, chunk + static_cast<uint64_t>(r) * nvrx::kCrcRowBytes + lane * 16, 16);  // the lane's uint4 load
            nvrx::crc_row_step(z512, t[lane], w[0], w[1], w[2], w[3]);
        }
    uint32_t s = 0;
    for (uint32_t lane = 0; lane < 32; ++lane) s = nvrx::crc_chain_lane(z16, s, nvrx::crc_fold_lane(z4, t[lane]));
    *out = s;
    return 0;
}

// Same walk through the lane-replicated table of crc_chunks_private<> (entry e of lane l at z512x32[e * 32 + l]).
extern "C" int crc_lanes_chunk_value_private(const uint8_t* chunk, uint32_t rows, const uint32_t* z512x32, const uint32_t* z4,
                                             const uint32_t* z16, uint32_t* out) {
    if (!chunk || !rows || rows > nvrx::kCrcChunkRows || !out) return 1;
    uint32_t t[32][4];
    memset(t, 0, sizeof(t));
    for (uint32_t r = 0; r < rows; ++r)
        for (uint32_t lane = 0; lane < 32; ++lane) {
            uint32_t w[4];
            memcpy(w, chunk + static_cast<uint64_t>(r) * nvrx::kCrcRowBytes + lane * 16, 16);
            nvrx::crc_row_step_lane(z512x32, lane, t[lane], w[0], w[1], w[2], w[3]);
        }
    uint32_t s = 0;
    for (uint32_t lane = 0; lane < 32; ++lane) s = nvrx::crc_chain_lane(z16, s, nvrx::crc_fold_lane(z4, t[lane]));
    *out = s;
    return 0;
}
