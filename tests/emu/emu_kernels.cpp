// tests/emu/emu_kernels.cpp -- compiles the CUDA kernel sources with g++ on top of the fiber
// SIMT emulator and exposes C entry points for pytest.  TEST INFRASTRUCTURE ONLY (see simt_emu.h).
#include "simt_emu.h"
#include "../../compress_b200/csrc/b2c_zstd_enc.cuh"
#include <vector>
#include <cstdlib>

using namespace b2c;

extern "C" {

void emu_set_lane_order(int desc) { emu::lane_order_desc = desc; }

// Encode nchunks chunks laid out contiguously (chunk i = src + i*stride, size sizes[i]).
// dst slots of dst_stride bytes.  Optional debug dumps (may be null).
int emu_zstd_encode(const uint8_t *src, uint64_t stride, const uint32_t *sizes, uint32_t nchunks, uint8_t *dst,
                    uint64_t dst_stride, int64_t *out_sizes, uint32_t flags, uint32_t *dbg_hdr, uint32_t *dbg_seqs,
                    uint8_t *dbg_lits, uint32_t dbg_seq_cap) {
    std::vector<uint8_t> scratch(ENC_SCRATCH_BYTES, 0xCD);
    ZstdEncParams P;
    memset(&P, 0, sizeof(P));
    P.srcs = nullptr; P.src_base = src; P.src_stride = stride; P.src_sizes = sizes; P.src_size_all = 0;
    P.dsts = nullptr; P.dst_base = dst; P.dst_stride = dst_stride; P.dst_cap = (uint32_t)dst_stride;
    P.out_sizes = out_sizes; P.nchunks = nchunks; P.flags = flags; P.scratch = scratch.data();
    P.dbg_hdr = dbg_hdr; P.dbg_seqs = dbg_seqs; P.dbg_lits = dbg_lits; P.dbg_seq_cap = dbg_seq_cap;
    emu::launch(1, ENC_NT, ENC_SMEM_BYTES, [&]() {
        uint8_t *smem = emu::dyn_smem;
        EncShared *sh = reinterpret_cast<EncShared *>(smem + ENC_SMEM_SH);
        if (threadIdx.x < 3) seq_build_predef(&sh->sw, (int)threadIdx.x);
        __syncthreads();
        for (uint32_t c = 0; c < P.nchunks; c++) zstd_encode_chunk(smem, P, c, P.scratch);
    });
    return 0;
}

uint32_t emu_enc_smem_bytes() { return ENC_SMEM_BYTES; }
}
