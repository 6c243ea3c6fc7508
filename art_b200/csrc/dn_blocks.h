// Detail recovery of RGB_denoise (FTblockDN.cc detail_recovery L1479-1635): the 64 x 64 block DCTs on tcgen05 (dn_blocks.cu).
#pragma once
#include "ctx.h"

struct DnBlocksArgs {
    const float* Lin;            // luminance before the wavelet stage, dense W x H
    const float* L;              // luminance after it
    const float* mask;           // detail mask (W x H) or nullptr
    int width, height, nbw, nbh;
    const float* tin;            // tilemask_in, 64 x 64 dense
    const unsigned* fwd_split;   // REDFT10 matrix, TF32 big | small halves in the canonical K-major operand layout (2 x 16 KB)
    const unsigned* bwd_split;   // REDFT01 matrix, the same
    float* blocks;               // [nbh][nbw][64][64]
    float detail_hi, detail_lo, params_Ldetail;
    int use_mask, blur_rad;
};

// bytes of the two pre-split operand images (forward | backward)
constexpr size_t DN_SPLIT_WORDS = 2 * 2 * 64 * 64;
// fills `host` (DN_SPLIT_WORDS words) from the dense forward / backward matrices [k][j]
void art_dn_blocks_split_tables(const float* dctf, const float* dctb, unsigned* host);
int art_dn_blocks_launch(art_hp_ctx* ctx, const DnBlocksArgs& a);
