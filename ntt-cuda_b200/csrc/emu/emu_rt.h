// emu_rt.h -- minimal CPU runtime for the kernel sources compiled with -DNTTB200_EMU: one std::thread per CUDA
// thread, std::barrier for __syncthreads()/__syncwarp(), and a functional model of TMA tile copies including the
// 128-byte shared-memory swizzle (address bits [4,7) ^= bits [7,10)).  TEST INFRASTRUCTURE: used only by tests/.
#pragma once
#include <barrier>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#include "../compat.cuh"

struct EmuTmapDesc {          // stored inside the opaque TensorMap bytes
    unsigned char *base;
    int rank;
    uint64_t dims[4];
    uint64_t strides[4];      // bytes; strides[0] = element size
    uint32_t box[4];
    int swizzle128;
};

struct EmuCta {
    std::unique_ptr<std::barrier<>> block_bar;
    std::vector<std::unique_ptr<std::barrier<>>> warp_bars;
};
extern thread_local EmuCta *emu_cta;

template <class F>
void emu_launch(emu_dim3 grid, unsigned block, size_t smem_bytes, F &&body)
{
    for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
        for (unsigned bx = 0; bx < grid.x; bx++) {
            std::vector<unsigned char> smem(smem_bytes + 16);
            EmuCta cta;
            cta.block_bar = std::make_unique<std::barrier<>>(block);
            for (unsigned w = 0; w < (block + 31) / 32; w++) {
                unsigned cnt = (w + 1) * 32 <= block ? 32 : block - w * 32;
                cta.warp_bars.push_back(std::make_unique<std::barrier<>>(cnt));
            }
            std::vector<std::thread> th;
            th.reserve(block);
            for (unsigned t = 0; t < block; t++)
                th.emplace_back([&, t] {
                    threadIdx.x = t; blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
                    blockDim.x = block; gridDim = grid;
                    emu_dyn_smem = smem.data();
                    emu_cta = &cta;
                    body();
                });
            for (auto &x : th) x.join();
        }
}
