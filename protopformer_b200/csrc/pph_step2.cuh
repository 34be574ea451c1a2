// Scratch layouts shared by the kernels of the fused training step (pph_prep.cu, pph_mid.cu, pph_simgrad2.cu).
#pragma once

#include "pph_bins.cuh"
#include "pph_common.cuh"

namespace pph {

// Token bins of one batch (written by the BIN role of pph_head_mid, read by pph_similarity_bwd2):
//   bin_start [B][K+1]  offsets into bin_list[b], bin k = entries [bin_start[k], bin_start[k+1])
//   bin_list  [B][P]    the image's prototypes sorted by argmin token (stable: ascending inside a bin)
//   item_start[B][K+1]  work-item offsets of the round-1 gradient kernel (unused by the new one)
//   cls_id [B], cls_start [n_cls+1], cls_item [n_cls+1], cls_order [B]: the images sorted by (clamped) label, stable --
//   the prototype-gradient kernel adds the per-image PPC rows of a class in image order
struct Step2Bins {
    int32_t *bin_start, *item_start, *bin_list;
    int32_t *cls_id, *cls_start, *cls_item, *cls_order;
    int4* item_desc;         // [B][bin_items_per_image(K, P)] work-item descriptors of the gather kernel (pph_bins.cuh)
    size_t bytes;
};

inline Step2Bins carve_bins(void* base, int B, int K, int P) {
    Step2Bins w;
    char* q = static_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t n) { char* r = q ? q + off : nullptr; off += (n + 255) / 256 * 256; return r; };
    w.bin_start = reinterpret_cast<int32_t*>(take(sizeof(int) * (size_t)B * (K + 1)));
    w.item_start = reinterpret_cast<int32_t*>(take(sizeof(int) * (size_t)B * (K + 1)));
    w.bin_list = reinterpret_cast<int32_t*>(take(sizeof(int) * (size_t)B * P));
    w.cls_id = reinterpret_cast<int32_t*>(take(sizeof(int) * (size_t)B));
    w.cls_start = reinterpret_cast<int32_t*>(take(sizeof(int) * (size_t)(P + 1)));      // n_cls = P / m <= P
    w.cls_item = reinterpret_cast<int32_t*>(take(sizeof(int) * (size_t)(P + 1)));
    w.cls_order = reinterpret_cast<int32_t*>(take(sizeof(int) * (size_t)B));
    w.item_desc = reinterpret_cast<int4*>(take(sizeof(int4) * (size_t)B * bin_items_per_image(K, P)));
    w.bytes = off + 256;
    return w;
}

}  // namespace pph
