// encode.cu -- K6/K7: entropy encode (statistics -> optimal Huffman tables -> bit packing -> byte stuffing).
// (work in progress: the entry point exists so that the C-ABI is complete; implemented below)
#include "common.cuh"

int jpeg_huffman_encode_scan(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan, const jpeg_sm100_dev_spectral *sp,
                             uint64_t interval_mcus, jpeg_sm100_huff_table *tables_out, uint8_t *d_ecs,
                             uint64_t ecs_image_stride, uint64_t *d_ecs_len, uint64_t *h_needed)
{
    (void) ctx; (void) scan; (void) sp; (void) interval_mcus; (void) tables_out; (void) d_ecs;
    (void) ecs_image_stride; (void) d_ecs_len; (void) h_needed;
    return JPEG_SM100_ERR_UNSUPPORTED;
}
