// sfmmatch.cu — implementation of the C ABI declared in include/sfmmatch.h.
//
// Host side of the matching stage: device bank management, batching of the pair list, kernel
// sequencing on one CUDA stream, result marshalling.  No CPU compute fallback exists: every entry
// point that computes distances launches the kernels of csrc/ or fails with SFM_ERR_CUDA.
#include "ctx_internal.h"

using namespace sfm;
using namespace sfmhost;

static_assert(sizeof(sfm_dmatch) == 16 && sizeof(DMatch) == 16, "DMatch must be byte-compatible with cv::DMatch");

namespace sfmhost {

std::string g_create_error;

int make_tmaps_f32(sfm_ctx* c, Bank& b) {
    b.f_tc_ok = false;
    if (b.padded_rows == 0) return SFM_OK;
    const cuuint64_t dims[2] = {512, static_cast<cuuint64_t>(b.padded_rows)};      // 128 floats = 512 bytes per row
    const cuuint64_t strides[1] = {512};
    const cuuint32_t estr[2] = {1, 1};
    const cuuint32_t box[2] = {128, 128};                                          // one 128-byte K slab x 128 rows
    void* src[4] = {b.d_fhi.p, b.d_flo.p, b.d_fhi.p, b.d_flo.p};
    for (int i = 0; i < 4; ++i) {
        CUresult r = c->encode(&b.tmaps_f[i], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, src[i], dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(c, SFM_ERR_CUDA, "cuTensorMapEncodeTiled(f32 hi/lo) failed: " + std::to_string(r));
    }
    const cuuint64_t dims_e[2] = {32, static_cast<cuuint64_t>(b.padded_rows)};
    const cuuint64_t strides_e[1] = {32};
    const cuuint32_t box_e[2] = {32, 128};
    CUresult r = c->encode(&b.tmaps_f[4], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, b.d_fext.p, dims_e, strides_e, box_e, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(c, SFM_ERR_CUDA, "cuTensorMapEncodeTiled(f32 norm rows) failed: " + std::to_string(r));
    b.f_tc_ok = true;
    return SFM_OK;
}

int make_tmaps(sfm_ctx* c, Bank& b) {
    b.have_tmap = false;
    if (b.padded_rows == 0) return SFM_OK;
    if (b.depth == SFM_CV_8U && b.cols == 32) {
        // ORB: tensor maps over the bit-expanded copy, 256-byte rows, one 128-byte-wide box per K slab
        const cuuint64_t dims[2] = {256, static_cast<cuuint64_t>(b.padded_rows)};
        const cuuint64_t strides[1] = {256};
        const cuuint32_t estr[2] = {1, 1};
        const cuuint32_t box_a[2] = {128, kTcRowsPerUnit};
        const cuuint32_t box_b[2] = {128, 256};
        CUresult r = c->encode(&b.tmap_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, b.d_bits.p, dims, strides, box_a, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS)
            r = c->encode(&b.tmap_b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, b.d_bits.p, dims, strides, box_b, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(c, SFM_ERR_CUDA, "cuTensorMapEncodeTiled(ORB bits) failed: " + std::to_string(r));
        b.have_tmap = true;
        return SFM_OK;
    }
    if (!(b.u8_valued && b.cols == 128)) return SFM_OK;
    const cuuint64_t dims[2] = {128, static_cast<cuuint64_t>(b.padded_rows)};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t estr[2] = {1, 1};
    const cuuint32_t box_a[2] = {128, kTcRowsPerUnit};
    const cuuint32_t box_b[2] = {128, 256};
    CUresult r = c->encode(&b.tmap_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, b.d_u8.p, dims, strides, box_a, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(c, SFM_ERR_CUDA, "cuTensorMapEncodeTiled(A) failed: " + std::to_string(r));
    r = c->encode(&b.tmap_b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, b.d_u8.p, dims, strides, box_b, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(c, SFM_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed: " + std::to_string(r));
    const cuuint64_t dims_e[2] = {kExtBytes, static_cast<cuuint64_t>(b.padded_rows)};
    const cuuint64_t strides_e[1] = {kExtBytes};
    const cuuint32_t box_e[2] = {kExtBytes, 256};
    r = c->encode(&b.tmap_e, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, b.d_ext.p, dims_e, strides_e, box_e, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(c, SFM_ERR_CUDA, "cuTensorMapEncodeTiled(E) failed: " + std::to_string(r));
    b.have_tmap = true;
    return SFM_OK;
}

// Lay out the bank (row offsets), allocate, zero the padding.
int bank_layout(sfm_ctx* c, Bank& b, int n_images, const int32_t* n_rows, int cols, int depth) {
    if (n_images < 0 || cols <= 0) return fail(c, SFM_ERR_INVALID, "bank: bad n_images/cols");
    if (depth != SFM_CV_8U && depth != SFM_CV_32F) return fail(c, SFM_ERR_INVALID, "bank: depth must be CV_8U or CV_32F");
    if (depth == SFM_CV_8U && cols % 16 != 0) return fail(c, SFM_ERR_INVALID, "bank: CV_8U descriptors need cols % 16 == 0");
    if (depth == SFM_CV_32F && cols % 4 != 0) return fail(c, SFM_ERR_INVALID, "bank: CV_32F descriptors need cols % 4 == 0");
    b.n_images = n_images; b.cols = cols; b.depth = depth; b.have_kp = false;
    b.n_rows.assign(n_rows, n_rows + n_images);
    b.row0.resize(n_images + 1);
    int64_t r = 0;
    for (int i = 0; i < n_images; ++i) {
        if (n_rows[i] < 0) return fail(c, SFM_ERR_INVALID, "bank: negative row count");
        if (n_rows[i] >= SFM_MAX_ROWS)
            return fail(c, SFM_ERR_CAPACITY, "bank: image has >= 2^18 descriptors (OpenCV IMGIDX_ONE limit)");
        b.row0[i] = r;
        r += pad_rows(n_rows[i]);
    }
    b.row0[n_images] = r;
    b.padded_rows = r;
    if (r >= (int64_t(1) << 31)) return fail(c, SFM_ERR_CAPACITY, "bank: more than 2^31 padded rows");
    return SFM_OK;
}

// After the raw descriptors are on the device (d_f32 for CV_32F, d_u8 for CV_8U): pack / norms / tensor maps.
int bank_finish(sfm_ctx* c, Bank& b) {
    NvtxRange nvtx_range("sfm:bank_finish (pack, norms, tensor maps)");
    cudaStream_t s = c->stream;
    b.u8_valued = false; b.have_f32 = false;
    if (b.padded_rows == 0) { b.u8_valued = b.depth == SFM_CV_8U; return make_tmaps(c, b); }
    // valid rows per 256-row block (pinned temporary: no sync needed before it is reused, see meta_ev)
    const int64_t nblk = b.padded_rows / kRowAlign;
    CU_TRY(c, cudaEventSynchronize(c->valid_ev));
    CU_TRY(c, c->h_valid.ensure(static_cast<size_t>(nblk) * 4));
    int32_t* valid = c->h_valid.as<int32_t>();
    for (int i = 0; i < b.n_images; ++i) {
        const int64_t b0 = b.row0[i] / kRowAlign, nb = pad_rows(b.n_rows[i]) / kRowAlign;
        for (int64_t k = 0; k < nb; ++k)
            valid[b0 + k] = static_cast<int32_t>(std::min<int64_t>(kRowAlign, std::max<int64_t>(0, b.n_rows[i] - k * kRowAlign)));
    }
    CU_TRY(c, b.d_valid.ensure(static_cast<size_t>(nblk) * 4));
    CU_TRY(c, cudaMemcpyAsync(b.d_valid.p, valid, static_cast<size_t>(nblk) * 4, cudaMemcpyHostToDevice, s));
    CU_TRY(c, cudaEventRecord(c->valid_ev, s));
    const int32_t* d_valid = b.d_valid.as<int32_t>();
    int* flags = reinterpret_cast<int*>(c->d_scalars.as<uint8_t>() + 16);      // [0] not-integer, [1] max |b|^2, [2] min |b|^2
    CU_TRY(c, cudaMemsetAsync(flags, 0, 8, s));
    CU_TRY(c, cudaMemsetAsync(flags + 2, 0x7f, 4, s));
    bool maybe_u8 = b.depth == SFM_CV_8U;
    if (b.depth == SFM_CV_32F && b.cols == 128) {
        CU_TRY(c, b.d_u8.ensure(static_cast<size_t>(b.padded_rows) * 128));
        CU_TRY(c, launch_pack_f32_to_u8(b.d_f32.as<float>(), 128, static_cast<int>(b.padded_rows), 128, d_valid,
                                        b.d_u8.as<uint8_t>(), flags, s));
        c->stat_launches++;
        maybe_u8 = true;
    } else if (b.depth == SFM_CV_8U) {
        CU_TRY(c, launch_zero_padding(b.d_u8.p, b.cols, b.padded_rows, d_valid, s));
        c->stat_launches++;
        if (b.cols == 32) {
            CU_TRY(c, b.d_bits.ensure(static_cast<size_t>(b.padded_rows) * 256));
            CU_TRY(c, b.d_norm2.ensure(b.padded_rows * 4));
            CU_TRY(c, b.d_ckey.ensure(b.padded_rows * 4));
            CU_TRY(c, launch_expand_bits(b.d_u8.as<uint8_t>(), b.padded_rows, d_valid, b.d_bits.as<uint8_t>(),
                                         b.d_norm2.as<int32_t>(), b.d_ckey.as<int32_t>(), s));
            c->stat_launches++;
        }
    }
    if (maybe_u8 && b.cols == 128) {
        CU_TRY(c, b.d_norm2.ensure(b.padded_rows * 4));
        CU_TRY(c, b.d_ckey.ensure(b.padded_rows * 4));
        CU_TRY(c, b.d_ext.ensure(static_cast<size_t>(b.padded_rows) * kExtBytes));
        const size_t nblk_b = static_cast<size_t>(b.padded_rows / kRowAlign) * 4;
        CU_TRY(c, b.d_blkmin.ensure(nblk_b));
        CU_TRY(c, b.d_blkmax.ensure(nblk_b));
        CU_TRY(c, cudaMemsetAsync(b.d_blkmin.p, 0x7f, nblk_b, s));
        CU_TRY(c, cudaMemsetAsync(b.d_blkmax.p, 0, nblk_b, s));
        CU_TRY(c, launch_norms_ckeys(b.d_u8.as<uint8_t>(), b.padded_rows, d_valid, b.d_norm2.as<int32_t>(),
                                     b.d_ckey.as<int32_t>(), b.d_ext.as<int8_t>(), flags + 1, b.d_blkmin.as<int32_t>(),
                                     b.d_blkmax.as<int32_t>(), s));
        c->stat_launches++;
    }
    int* h = c->h_scalars.as<int>() + 8;
    CU_TRY(c, cudaMemcpyAsync(h, flags, 12, cudaMemcpyDeviceToHost, s));
    CU_TRY(c, cudaStreamSynchronize(s));
    b.ext_ok = false;
    if (b.depth == SFM_CV_32F) {
        // integer-valued (what cv::SIFT emits): match on the u8 copy; d_f32 stays allocated as upload staging
        if (b.cols == 128 && h[0] == 0) { b.u8_valued = true; b.have_f32 = false; }
        else {
            b.have_f32 = true;
            CU_TRY(c, launch_zero_padding(b.d_f32.p, b.cols * 4, b.padded_rows, d_valid, s));
            c->stat_launches++;
            b.f_tc_ok = false;
            if (b.cols == 128) {
                // hi/lo split, norms and norm rows for the 3xTF32 tcgen05 kernel
                const size_t fb = static_cast<size_t>(b.padded_rows) * 512;
                CU_TRY(c, b.d_fhi.ensure(fb));
                CU_TRY(c, b.d_flo.ensure(fb));
                CU_TRY(c, b.d_fnorm.ensure(b.padded_rows * 4));
                CU_TRY(c, b.d_fext.ensure(static_cast<size_t>(b.padded_rows) * 32));
                CU_TRY(c, cudaMemsetAsync(flags, 0, 8, s));
                CU_TRY(c, launch_f32_split(b.d_f32.as<float>(), b.padded_rows, d_valid, b.d_fhi.as<float>(), b.d_flo.as<float>(),
                                           b.d_fnorm.as<float>(), b.d_fext.as<float>(), flags + 1, s));
                c->stat_launches++;
                CU_TRY(c, cudaMemcpyAsync(h, flags, 8, cudaMemcpyDeviceToHost, s));
                CU_TRY(c, cudaStreamSynchronize(s));
                std::memcpy(&b.f_nb_max, &h[1], 4);
                int rcf = make_tmaps_f32(c, b);
                if (rcf != SFM_OK) return rcf;
            }
        }
    } else {
        b.u8_valued = true;
    }
    if (b.u8_valued && b.cols == 128) {
        b.ext_ok = h[1] <= kExtMaxNorm2; b.nb_max = h[1]; b.nb_min = std::min(h[2], h[1]);
        if (&b == &c->bank) { c->prev_nb_min = b.nb_min; c->prev_nb_max = b.nb_max; }
    }
    return make_tmaps(c, b);
}

int bank_upload_host(sfm_ctx* c, Bank& b, int n_images, const void* const* rows, const int32_t* n_rows, int cols,
                     const size_t* step_bytes, int depth) {
    int rc = bank_layout(c, b, n_images, n_rows, cols, depth);
    if (rc != SFM_OK) return rc;
    cudaStream_t s = c->stream;
    const size_t esz = depth == SFM_CV_32F ? 4 : 1;
    const size_t row_bytes = static_cast<size_t>(cols) * esz;
    DevBuf& dst = depth == SFM_CV_32F ? b.d_f32 : b.d_u8;
    CU_TRY(c, dst.ensure(std::max<size_t>(16, static_cast<size_t>(b.padded_rows) * row_bytes)));
    // Caller buffers that are already page-locked (cudaHostAlloc / cudaHostRegister, e.g. torch pinned tensors) are
    // DMA'd directly; pageable ones are staged through two pinned buffers so that the H2D copy of one chunk
    // overlaps the host gather of the next.
    const size_t chunk = size_t(32) << 20;
    int which = 0;
    for (int i = 0; i < n_images; ++i) {
        if (n_rows[i] == 0) continue;
        if (!rows[i]) return fail(c, SFM_ERR_INVALID, "bank: null descriptor pointer for a non-empty image");
        const size_t step = step_bytes ? step_bytes[i] : row_bytes;
        if (step < row_bytes) return fail(c, SFM_ERR_INVALID, "bank: step smaller than a row");
        uint8_t* d_img = static_cast<uint8_t*>(dst.p) + static_cast<size_t>(b.row0[i]) * row_bytes;
        cudaPointerAttributes attr;
        const bool pinned = cudaPointerGetAttributes(&attr, rows[i]) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        if (!pinned) cudaGetLastError();
        if (pinned) {
            if (step == row_bytes)
                CU_TRY(c, cudaMemcpyAsync(d_img, rows[i], static_cast<size_t>(n_rows[i]) * row_bytes, cudaMemcpyHostToDevice, s));
            else
                CU_TRY(c, cudaMemcpy2DAsync(d_img, row_bytes, rows[i], step, row_bytes, static_cast<size_t>(n_rows[i]),
                                            cudaMemcpyHostToDevice, s));
            c->stat_h2d += static_cast<int64_t>(n_rows[i]) * static_cast<int64_t>(row_bytes);
            continue;
        }
        for (int k = 0; k < 2; ++k) CU_TRY(c, c->h_stage[k].ensure(chunk));
        const size_t rows_per_chunk = std::max<size_t>(1, chunk / row_bytes);
        for (size_t r0 = 0; r0 < static_cast<size_t>(n_rows[i]); r0 += rows_per_chunk) {
            const size_t nr = std::min(rows_per_chunk, static_cast<size_t>(n_rows[i]) - r0);
            CU_TRY(c, cudaEventSynchronize(c->stage_ev[which]));
            uint8_t* h = c->h_stage[which].as<uint8_t>();
            const uint8_t* src = static_cast<const uint8_t*>(rows[i]) + r0 * step;
            if (step == row_bytes) std::memcpy(h, src, nr * row_bytes);
            else for (size_t r = 0; r < nr; ++r) std::memcpy(h + r * row_bytes, src + r * step, row_bytes);
            CU_TRY(c, cudaMemcpyAsync(d_img + r0 * row_bytes, h, nr * row_bytes, cudaMemcpyHostToDevice, s));
            CU_TRY(c, cudaEventRecord(c->stage_ev[which], s));
            c->stat_h2d += static_cast<int64_t>(nr * row_bytes);
            which ^= 1;
        }
    }
    return bank_finish(c, b);
}

// ------------------------------------------------------------------------------------------------ the stage
enum class Engine { TC, TCV, TCN, TF32, DP4A, F32, POPC };

int pick_engine(sfm_ctx* c, const Bank& b, int norm, int requested, Engine* out) {
    if (norm == SFM_NORM_HAMMING) {
        if (b.depth != SFM_CV_8U) return fail(c, SFM_ERR_INVALID, "NORM_HAMMING needs CV_8U descriptors");
        if (b.cols != 32) return fail(c, SFM_ERR_UNSUPPORTED, "Hamming kernel is built for 256-bit (32-byte) descriptors");
        // tensor engine: bits expanded to bytes, same tcgen05 kernel with K = 256 (exact); SIMT: the __popc kernel
        *out = requested == SFM_ENGINE_SIMT ? Engine::POPC : Engine::TC;
        return SFM_OK;
    }
    if (norm != SFM_NORM_L2) return fail(c, SFM_ERR_UNSUPPORTED, "only NORM_L2 and NORM_HAMMING are supported");
    if (b.u8_valued && b.cols == 128) {
        *out = requested == SFM_ENGINE_SIMT ? Engine::DP4A : Engine::TC;
        return SFM_OK;
    }
    if (!b.have_f32) return fail(c, SFM_ERR_UNSUPPORTED, "NORM_L2 on CV_8U data needs 128-byte descriptors");
    // non-integer float descriptors: 3xTF32 tcgen05 candidate search + exact fp32 re-rank (128 columns), or the
    // CUDA-core fp32 kernel (any width up to 512, and SFM_ENGINE_SIMT)
    if (b.f_tc_ok && requested != SFM_ENGINE_SIMT) { *out = Engine::TF32; return SFM_OK; }
    if (requested == SFM_ENGINE_TENSOR || requested == SFM_ENGINE_TENSOR_IMAD)
        return fail(c, SFM_ERR_UNSUPPORTED, "tensor engine needs 128-column descriptors");
    if (b.cols > 512) return fail(c, SFM_ERR_UNSUPPORTED, "fp32 L2 kernel supports up to 512 columns");
    *out = Engine::F32;
    return SFM_OK;
}

int launch_knn(sfm_ctx* c, const Bank& b, Engine eng, const PairDesc* d_pairs, const int64_t* d_unit_prefix, int n_pairs,
               int64_t n_units, Top2* out, float* aux = nullptr, const TcvFuse* fz = nullptr, int grid_limit = 0) {
    cudaStream_t s = c->stream;
    c->stat_launches++;
    const TcvFuse none{};
    const int sms = grid_limit > 0 ? std::min(grid_limit, c->sm_count) : c->sm_count;
    switch (eng) {
        case Engine::TC:
            CU_TRY(c, launch_knn2_l2_u8_tc(&b.tmap_a, &b.tmap_b, b.d_ckey.as<int32_t>(), b.d_norm2.as<int32_t>(), d_pairs,
                                           d_unit_prefix, n_pairs, n_units, out, c->sm_count, b.cols == 32 ? 2 : 1, s));
            break;
        case Engine::TCV:
            CU_TRY(c, launch_knn2_l2_u8_tcv(&b.tmap_a, &b.tmap_b, &b.tmap_e, d_pairs, d_unit_prefix, n_pairs, n_units, out,
                                            nullptr, sms, c->tcv_layout_run, 256, c->tcv_issuers, c->tcv_chunk_run, fz ? *fz : none, s));
            break;
        case Engine::TCN:
            CU_TRY(c, launch_knn2_l2_u8_tcv(&b.tmap_a, &b.tmap_b, &b.tmap_e, d_pairs, d_unit_prefix, n_pairs, n_units, out,
                                            reinterpret_cast<int32_t*>(aux), sms, c->tcv_layout_run, 256, c->tcv_issuers,
                                            c->tcv_chunk_run, fz ? *fz : none, s));
            break;
        case Engine::TF32:
            CU_TRY(c, launch_knn2_l2_f32_tc3(b.tmaps_f, d_pairs, d_unit_prefix, n_pairs, n_units, out, aux, c->sm_count, s));
            break;
        case Engine::DP4A:
            CU_TRY(c, launch_knn2_l2_u8_dp4a(b.d_u8.as<uint8_t>(), b.d_norm2.as<int32_t>(), d_pairs, d_unit_prefix, n_pairs,
                                             n_units, out, s));
            break;
        case Engine::F32:
            CU_TRY(c, launch_knn2_l2_f32(b.d_f32.as<float>(), b.cols, d_pairs, d_unit_prefix, n_pairs, n_units, out, s));
            break;
        case Engine::POPC:
            CU_TRY(c, launch_knn2_hamming_popc(b.d_u8.as<uint8_t>(), d_pairs, d_unit_prefix, n_pairs, n_units, out, s));
            break;
    }
    return SFM_OK;
}

int rows_per_unit(Engine e) {
    return e == Engine::F32 ? kF32RowsPerUnit
                            : ((e == Engine::TC || e == Engine::TCV || e == Engine::TCN || e == Engine::TF32) ? kTcRowsPerUnit : kSimtRowsPerUnit);
}

int enqueue_impl(sfm_ctx* c, const int32_t* pairs_in, int64_t n_pairs, const sfm_opts* o, const Schedule* sched) {
    NvtxRange nvtx_range("sfm:match_pairs_enqueue");
    Bank& b = c->bank;
    const int32_t* pairs = pairs_in;
    std::vector<int32_t> scheduled;
    if (sched && sched->order && n_pairs > 0) {
        if (!pairs_in) return fail(c, SFM_ERR_INVALID, "bad pair list");
        scheduled.resize(2 * n_pairs);
        for (int64_t k = 0; k < n_pairs; ++k) {
            scheduled[2 * k] = pairs_in[2 * sched->order[k]];
            scheduled[2 * k + 1] = pairs_in[2 * sched->order[k] + 1];
        }
        pairs = scheduled.data();           // from here on "pair p" means schedule position p
    }
    if (b.n_images == 0 && n_pairs > 0) return fail(c, SFM_ERR_STATE, "match_pairs before bank upload");
    if (n_pairs < 0 || (n_pairs > 0 && !pairs)) return fail(c, SFM_ERR_INVALID, "bad pair list");
    if (n_pairs >= (int64_t(1) << 31)) return fail(c, SFM_ERR_CAPACITY, "too many pairs");
    if (o->k != 1 && o->k != 2) return fail(c, SFM_ERR_UNSUPPORTED, "k must be 1 or 2");
    if (!(o->ratio >= 0.0)) return fail(c, SFM_ERR_INVALID, "ratio must be >= 0");
    Engine eng;
    int rc = pick_engine(c, b, o->norm, o->engine, &eng);
    if (rc != SFM_OK) return rc;
    // ratio-test runs on data with SIFT-sized norms take the value-only tcgen05 kernel (knn_l2_tcv.cu)
    // (its 32-bit running keys number at most 512 chunks per epilogue warp: train images <= 65536 rows)
    if (eng == Engine::TC && o->engine != SFM_ENGINE_TENSOR_IMAD && o->k == 2 && !o->cross_check && b.ext_ok) {
        int32_t max_rows = 0;
        for (int32_t n : b.n_rows) max_rows = std::max(max_rows, n);
        // 2 epilogue groups (8 warps, layout 12) measured fastest; their 32-bit keys cover train images up to 32768
        // rows, 4 groups (layout 14) up to 65536
        c->tcv_layout_run = c->tcv_layout ? c->tcv_layout : (max_rows <= 32768 ? 12 : 14);
        if (max_rows <= (c->tcv_layout_run == 14 ? 65536 : 32768)) {
            eng = Engine::TCV;
            // feedback of the previous run (copied to pinned memory behind its kernels; never waited for)
            if (c->tune_pending && cudaEventQuery(c->tune_ev) == cudaSuccess) {
                const double frac = c->tune_rows > 0 ? static_cast<double>(c->h_tune.as<unsigned long long>()[0]) / static_cast<double>(c->tune_rows) : 0.0;
                if (frac > 0.02) c->dense_matches = true;
                else if (frac < 0.01) c->dense_matches = false;
                c->tune_pending = false;
            }
            c->tcv_chunk_run = c->tcv_chunk ? c->tcv_chunk : (c->dense_matches ? 32 : 64);
            if (c->tcv_chunk_run == 128) c->tcv_chunk_run = 64;     // 128-row chunks exist for the in-kernel re-rank variant only (set below)
            // norms within 1/2 of each other (cv::SIFT: ~1 %, the SURVEY 8d recipe: 20-36 %): the norm-less variant, one K-step
            // less per tile; its bounds lose their grip when norms vary a lot, the exactness does not depend on the choice
            // (round 2: also for dense lists — with the re-rank inside the kernel the norm-less variant with 32-row chunks beats
            // the norm K-step there too: C4 538 k vs 488 k pairs/s, profiles/r2_epilogue_experiments.txt)
            if (c->tcv_normless == 2 || (c->tcv_normless == 1 &&
                                         static_cast<int64_t>(b.nb_max - b.nb_min) * c->tcv_spread_div <= b.nb_max))
                eng = Engine::TCN;
            // sparse lists with the in-kernel re-rank: 128-row chunks (half the per-chunk bookkeeping of the epilogue, twice the rows per
            // re-ranked chunk: +1.7 % on C3, profiles/r2_epilogue_experiments.txt)
            if (eng == Engine::TCN && (c->tcv_chunk == 128 || (c->tcv_chunk == 0 && !c->dense_matches)) && c->tcv_layout_run == 12 &&
                c->tcv_inkernel_refine)
                c->tcv_chunk_run = 128;
        }
    }
    for (int64_t p = 0; p < n_pairs; ++p) {
        const int l = pairs[2 * p], r = pairs[2 * p + 1];
        if (l < 0 || r < 0 || l >= b.n_images || r >= b.n_images) return fail(c, SFM_ERR_INVALID, "pair index out of range");
    }
    cudaStream_t s = c->stream;
    const int rpu = rows_per_unit(eng);
    const bool need_rev = o->cross_check != 0;
    const bool need_cnt = o->distinct != 0;

    // ---- split into batches bounded by the staging budget
    struct Batch { int64_t p0, p1, staged_rows, t_rows, n_units, n_rev_units; };
    std::vector<Batch> batches;
    {
        // the post kernels of batch i run beside the knn kernel of batch i + 1 (second stream, two buffer sets): a long list is
        // cut into at least c->min_batches pieces so that only the last piece's post work is exposed; short lists stay whole
        int64_t all_rows = 0;
        for (int64_t p = 0; p < n_pairs; ++p) all_rows += pad_rows(b.n_rows[pairs[2 * p]]);
        const int64_t budget = std::min<int64_t>(static_cast<int64_t>(c->staging_budget_rows),
                                                 std::max<int64_t>(all_rows / std::max(1, c->min_batches) + 1, int64_t(1) << 20));
        Batch cur{0, 0, 0, 0, 0, 0};
        for (int64_t p = 0; p < n_pairs; ++p) {
            const int64_t q = pad_rows(b.n_rows[pairs[2 * p]]), t = pad_rows(b.n_rows[pairs[2 * p + 1]]);
            const bool avail_break = sched && sched->avail && p > 0 && sched->avail[p] != sched->avail[p - 1];
            if (cur.p1 > cur.p0 && (avail_break || cur.staged_rows + q > budget || cur.t_rows + t > budget)) {
                batches.push_back(cur);
                cur = Batch{p, p, 0, 0, 0, 0};
            }
            cur.p1 = p + 1;
            cur.staged_rows += q;
            cur.t_rows += t;
        }
        if (cur.p1 > cur.p0) batches.push_back(cur);
    }
    const int64_t nb = static_cast<int64_t>(batches.size());

    // ---- host metadata for the whole list, one H2D
    //   PairDesc[n] (input order: filters) | PairDesc[n] (processing order: knn) | rev PairDesc[n] (processing order)
    //   | unit_prefix[n+nb] | rev_unit_prefix[..] | out_prefix[..] | t_prefix[..]
    // Processing order: inside a batch the pairs are walked in 8 x 8 image blocks, so that the ~16 descriptor
    // sets a block touches stay L2-resident (the input order of an all-pairs list re-reads every train image
    // from HBM for every pair).  Results still land at the staging rows of the INPUT order.
    const size_t npre = static_cast<size_t>(n_pairs + nb);
    const size_t bytes_pd = sizeof(PairDesc) * static_cast<size_t>(n_pairs);
    const size_t bytes_pre = sizeof(int64_t) * npre;
    const size_t bytes_meta = 3 * bytes_pd + 4 * bytes_pre;
    CU_TRY(c, cudaEventSynchronize(c->meta_ev));
    CU_TRY(c, c->h_meta.ensure(std::max<size_t>(64, bytes_meta)));
    PairDesc* h_pd = c->h_meta.as<PairDesc>();
    PairDesc* h_ppd = h_pd + n_pairs;
    PairDesc* h_rpd = h_ppd + n_pairs;
    int64_t* h_unit = reinterpret_cast<int64_t*>(h_rpd + n_pairs);
    int64_t* h_runit = h_unit + npre;
    int64_t* h_out = h_runit + npre;
    int64_t* h_tp = h_out + npre;
    int64_t max_staged = 0, max_t = 0, max_chunks = 0, total_q = 0;
    std::vector<int64_t> order;
    for (int64_t bi = 0; bi < nb; ++bi) {
        Batch& B = batches[bi];
        int64_t orow = 0, trow = 0;
        const int64_t base = B.p0 + bi;      // prefix arrays carry one extra entry per batch
        const int64_t np = B.p1 - B.p0;
        for (int64_t p = B.p0; p < B.p1; ++p) {
            const int l = pairs[2 * p], r = pairs[2 * p + 1];
            PairDesc d;
            d.q_row0 = static_cast<int32_t>(b.row0[l]); d.nq = b.n_rows[l];
            d.t_row0 = static_cast<int32_t>(b.row0[r]); d.nt = b.n_rows[r];
            d.out_row0 = orow;
            h_pd[p] = d;
            const int64_t k = p - B.p0;
            h_out[base + k] = orow; h_tp[base + k] = trow;
            orow += pad_rows(d.nq);
            trow += pad_rows(d.nt);
            total_q += d.nq;
        }
        h_out[base + np] = orow; h_tp[base + np] = trow;
        order.resize(np);
        for (int64_t k = 0; k < np; ++k) order[k] = B.p0 + k;
        std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) {
            const int lx = pairs[2 * x] >> 3, rx = pairs[2 * x + 1] >> 3, ly = pairs[2 * y] >> 3, ry = pairs[2 * y + 1] >> 3;
            return lx != ly ? lx < ly : rx < ry;
        });
        int64_t units = 0, runits = 0;
        for (int64_t k = 0; k < np; ++k) {
            const PairDesc& d = h_pd[order[k]];
            h_ppd[B.p0 + k] = d;
            PairDesc rd;                      // roles swapped: train rows query the left image
            rd.q_row0 = d.t_row0; rd.nq = d.nt; rd.t_row0 = d.q_row0; rd.nt = d.nq;
            rd.out_row0 = h_tp[base + (order[k] - B.p0)];
            h_rpd[B.p0 + k] = rd;
            h_unit[base + k] = units; h_runit[base + k] = runits;
            units += (d.nq + rpu - 1) / rpu;
            runits += (d.nt + rpu - 1) / rpu;
        }
        h_unit[base + np] = units; h_runit[base + np] = runits;
        B.n_units = units; B.n_rev_units = runits; B.staged_rows = orow; B.t_rows = trow;
        max_staged = std::max(max_staged, orow);
        max_t = std::max(max_t, trow);
        max_chunks = std::max(max_chunks, orow / 256);
    }
    CU_TRY(c, c->d_pairs.ensure(std::max<size_t>(64, bytes_meta)));
    if (n_pairs > 0) {
        CU_TRY(c, cudaMemcpyAsync(c->d_pairs.p, c->h_meta.p, bytes_meta, cudaMemcpyHostToDevice, s));
        CU_TRY(c, cudaEventRecord(c->meta_ev, s));
        c->stat_h2d += static_cast<int64_t>(bytes_meta);
    }
    const PairDesc* d_pd = c->d_pairs.as<PairDesc>();
    const PairDesc* d_ppd = d_pd + n_pairs;
    const PairDesc* d_rpd = d_ppd + n_pairs;
    const int64_t* d_unit = reinterpret_cast<const int64_t*>(d_rpd + n_pairs);
    const int64_t* d_runit = d_unit + npre;
    const int64_t* d_outp = d_runit + npre;
    const int64_t* d_tp = d_outp + npre;

    const bool sparse = eng == Engine::TCV || eng == Engine::TCN;       // fused ratio bound: need list + keep bits
    const int n_slots = nb > 1 ? 2 : 1;
    for (int k = 0; k < n_slots; ++k) {
        BatchSlot& S = c->slot[k];
        CU_TRY(c, S.top2.ensure(std::max<size_t>(16, static_cast<size_t>(max_staged) * sizeof(Top2))));
        if (need_rev) CU_TRY(c, S.rev.ensure(std::max<size_t>(16, static_cast<size_t>(max_t) * sizeof(Top2))));
        if (sparse) {
            CU_TRY(c, S.bf.ensure(std::max<size_t>(16, static_cast<size_t>(max_staged) * 4)));
            CU_TRY(c, S.need.ensure(std::max<size_t>(16, static_cast<size_t>(max_staged) * 4)));
            CU_TRY(c, S.done.ensure(std::max<size_t>(16, static_cast<size_t>(max_staged) * 4)));
            CU_TRY(c, S.keep.ensure(std::max<size_t>(32, static_cast<size_t>(max_staged) / 8)));
            CU_TRY(c, S.pair_nb.ensure(std::max<size_t>(16, static_cast<size_t>(n_pairs) * 8)));
        }
        CU_TRY(c, S.counters.ensure(32));
        if (eng == Engine::TF32 || eng == Engine::TCN) {
            CU_TRY(c, S.aux.ensure(std::max<size_t>(16, static_cast<size_t>(max_staged) * 4)));
            if (need_rev) CU_TRY(c, S.aux_rev.ensure(std::max<size_t>(16, static_cast<size_t>(max_t) * 4)));
        }
        if (need_cnt) CU_TRY(c, S.train_cnt.ensure(std::max<size_t>(16, static_cast<size_t>(max_t) * 4)));
        CU_TRY(c, S.chunk_counts.ensure(std::max<size_t>(16, static_cast<size_t>(max_chunks) * 4)));
        CU_TRY(c, S.blk_pair.ensure(std::max<size_t>(16, static_cast<size_t>(max_chunks) * 4)));
        CU_TRY(c, S.chunk_excl.ensure(static_cast<size_t>(max_chunks + 1) * 8));
        CU_TRY(c, S.pair_counts.ensure(std::max<size_t>(16, static_cast<size_t>(n_pairs) * 8)));
    }
    if (sparse || eng == Engine::TF32) CU_TRY(c, cudaMemsetAsync(c->d_scalars.as<uint8_t>() + 32, 0, 16, s));   // stats of this run
    CU_TRY(c, c->d_pair_offsets.ensure(std::max<size_t>(16, static_cast<size_t>(n_pairs) * 8)));
    CU_TRY(c, c->d_dropped.ensure(std::max<size_t>(16, static_cast<size_t>(n_pairs))));
    // output capacity: grown on demand (collect() retries with the worst case if a run overflows)
    if (c->out_capacity == 0) c->out_capacity = int64_t(1) << 22;
    c->out_capacity = std::max<int64_t>(c->out_capacity, std::min<int64_t>(total_q, std::max<int64_t>(int64_t(1) << 22, total_q / 4)));
    CU_TRY(c, c->d_out.ensure(static_cast<size_t>(c->out_capacity) * sizeof(DMatch)));
    CU_TRY(c, cudaMemsetAsync(c->d_scalars.p, 0, 12, s));     // running_total, overflow
    int64_t* d_total = c->d_scalars.as<int64_t>();
    int* d_overflow = reinterpret_cast<int*>(c->d_scalars.as<uint8_t>() + 8);

    FilterParams fp;
    fp.norm = o->norm; fp.k = o->k; fp.ratio = o->ratio; fp.cross_check = o->cross_check; fp.distinct = o->distinct;
    c->prof_used = 0;
    if (c->profiling) {
        while (static_cast<int64_t>(c->prof_ev.size()) < 3 * nb) {
            cudaEvent_t ev;
            CU_TRY(c, cudaEventCreate(&ev));
            c->prof_ev.push_back(ev);
        }
    }
    cudaStream_t ps = c->post_stream;
    for (int64_t bi = 0; bi < nb; ++bi) {
        const Batch& B = batches[bi];
        BatchSlot& S = c->slot[bi & 1];
        const int np = static_cast<int>(B.p1 - B.p0);
        const int64_t base = B.p0 + bi;
        // ---- compute stream: the knn kernel(s) of this batch.  The slot's buffers are free once the post kernels of batch
        // bi - 2 are done.
        if (bi >= 2) CU_TRY(c, cudaStreamWaitEvent(s, S.post_done, 0));
        if (sched && sched->avail && sched->events) CU_TRY(c, cudaStreamWaitEvent(s, sched->events[sched->avail[B.p0]], 0));
        CU_TRY(c, launch_block_pairs(d_outp + base, np, B.staged_rows / 256, S.blk_pair.as<int32_t>(), s));
        c->stat_launches++;
        TcvFuse fz{};
        if (sparse) {
            // all pairs of the batch with the same number of 128-row units (equal-size images): the kernels divide instead of searching
            const int64_t u0 = h_unit[base + 1] - h_unit[base];
            bool uniform = u0 > 0 && u0 < (int64_t(1) << 30);
            for (int k = 1; k < np && uniform; ++k) uniform = h_unit[base + k + 1] - h_unit[base + k] == u0;
            fz.uniform_units = uniform ? static_cast<int>(u0) : 0;
            CU_TRY(c, cudaMemsetAsync(S.counters.p, 0, 32, s));                            // brute-force queue, need list, done list lengths
            CU_TRY(c, cudaMemsetAsync(S.keep.p, 0, static_cast<size_t>(B.staged_rows) / 8, s));
            fz.norm2 = b.d_norm2.as<int32_t>(); fz.blk_min = b.d_blkmin.as<int32_t>(); fz.blk_max = b.d_blkmax.as<int32_t>();
            fz.need_list = S.need.as<int32_t>(); fz.need_count = S.counters.as<int>() + 1; fz.ratio = o->ratio;
            // norm-less variant with 8 epilogue warps: four more warps of every CTA re-rank the survivors in the kernel
            // dense lists (grid neighbours: ~7 % of the rows survive the bound): 2 = the epilogue waits for the finishing warps instead
            // of diverting their backlog to the post pass — a re-rank under the MMA costs ~0.78 of one in the post pass (C4, measured)
            fz.refine = (eng == Engine::TCN && c->tcv_layout_run == 12 && c->tcv_inkernel_refine) ? (c->tcv_divert_test ? 3 : (c->tcv_backpressure == 2 || (c->dense_matches && c->tcv_backpressure)) ? 2 : 1) : 0;
            fz.bank = b.d_u8.as<uint8_t>(); fz.done_list = S.done.as<int32_t>(); fz.done_count = S.counters.as<int>() + 2;
            fz.bf_list = S.bf.as<int32_t>(); fz.bf_count = S.counters.as<int>();
            fz.stats = reinterpret_cast<unsigned long long*>(c->d_scalars.as<uint8_t>() + 32);
        }
        if (c->profiling) CU_TRY(c, cudaEventRecord(c->prof_ev[3 * bi], s));
        // batches that run while later chunks of the bank are still being exchanged leave SMs to the NCCL kernels (dist from-host path)
        const bool chunks_pending = sched && sched->avail && sched->n_groups > 0 && sched->avail[B.p0] + 1 < sched->n_groups;
        rc = launch_knn(c, b, eng, d_ppd + B.p0, d_unit + base, np, B.n_units, S.top2.as<Top2>(), S.aux.as<float>(), &fz,
                        chunks_pending ? c->knn_grid_limit : 0);
        if (rc != SFM_OK) return rc;
        if (need_rev) {
            rc = launch_knn(c, b, eng, d_rpd + B.p0, d_runit + base, np, B.n_rev_units, S.rev.as<Top2>(), S.aux_rev.as<float>());
            if (rc != SFM_OK) return rc;
        }
        if (c->profiling) CU_TRY(c, cudaEventRecord(c->prof_ev[3 * bi + 1], s));
        CU_TRY(c, cudaEventRecord(S.knn_done, s));
        // ---- post stream: exact re-rank, filters, scan, ordered compaction (in batch order: the scan continues a running total)
        CU_TRY(c, cudaStreamWaitEvent(ps, S.knn_done, 0));
        if (eng == Engine::TF32) {
            // candidates -> exact fp32 top-2 (+ certificate); everything downstream sees ordinary Top2 rows
            RefineF32Args fa;
            fa.blk_pair = S.blk_pair.as<int32_t>();
            fa.top2 = S.top2.as<Top2>(); fa.aux = S.aux.as<float>(); fa.pairs = d_pd + B.p0; fa.out_prefix = d_outp + base;
            fa.n_pairs = np; fa.staged_rows = B.staged_rows; fa.bank = b.d_f32.as<float>(); fa.fnorm2 = b.d_fnorm.as<float>();
            fa.nb_max = b.f_nb_max; fa.all_rows = (o->k == 1 || need_rev) ? 1 : 0; fa.swap_roles = 0; fa.ratio = o->ratio;
            fa.stats = reinterpret_cast<unsigned long long*>(c->d_scalars.as<uint8_t>() + 32);
            CU_TRY(c, launch_refine_f32(fa, ps));
            c->stat_launches++;
            if (need_rev) {
                fa.top2 = S.rev.as<Top2>(); fa.aux = S.aux_rev.as<float>(); fa.swap_roles = 1; fa.blk_pair = nullptr;
                fa.out_prefix = d_tp + base; fa.staged_rows = B.t_rows; fa.all_rows = 1;
                CU_TRY(c, launch_refine_f32(fa, ps));
                c->stat_launches++;
            }
        }
        if ((eng == Engine::TC || sparse) && o->k == 2 && !need_rev) {
            RefineArgs ra{};
            ra.aux = S.aux.as<int32_t>(); ra.blk_min = b.d_blkmin.as<int32_t>(); ra.blk_max = b.d_blkmax.as<int32_t>();
            ra.stats = reinterpret_cast<unsigned long long*>(c->d_scalars.as<uint8_t>() + 32);
            ra.chunk_rows = c->tcv_chunk_run; ra.blk_pair = S.blk_pair.as<int32_t>();
            ra.bf_list = S.bf.as<int32_t>(); ra.bf_count = S.counters.as<int>();
            ra.need_list = S.need.as<int32_t>(); ra.need_count = ra.bf_count + 1; ra.pair_nb = S.pair_nb.as<int32_t>();
            ra.top2 = S.top2.as<Top2>(); ra.pairs = d_pd + B.p0; ra.out_prefix = d_outp + base; ra.n_pairs = np;
            ra.staged_rows = B.staged_rows; ra.bank = b.d_u8.as<uint8_t>(); ra.norm2 = b.d_norm2.as<int32_t>();
            ra.all_rows = 0; ra.ratio = o->ratio; ra.hamming = o->norm == SFM_NORM_HAMMING;
            if (eng == Engine::TCN) {
                CU_TRY(c, launch_refine_dot(ra, ps));
                c->stat_launches += 2;                                      // 3 kernels
            }
            else if (eng == Engine::TCV) CU_TRY(c, launch_refine_value(ra, ps));
            else CU_TRY(c, launch_refine_second(ra, ps));
            c->stat_launches++;
        }
        FilterArgs a{};
        a.top2 = S.top2.as<Top2>(); a.rev = S.rev.as<Top2>(); a.pairs = d_pd + B.p0;
        a.out_prefix = d_outp + base; a.t_prefix = d_tp + base; a.n_pairs = np; a.staged_rows = B.staged_rows;
        a.fp = fp; a.train_cnt = S.train_cnt.as<int32_t>(); a.blk_pair = S.blk_pair.as<int32_t>();
        if (need_cnt) CU_TRY(c, cudaMemsetAsync(S.train_cnt.p, 0, static_cast<size_t>(B.t_rows) * 4, ps));
        if (sparse) {
            a.keep_bits = S.keep.as<uint32_t>(); a.need_list = S.need.as<int32_t>(); a.need_count = S.counters.as<int>() + 1;
            a.done_list = S.done.as<int32_t>(); a.done_count = S.counters.as<int>() + 2;
            if (eng == Engine::TCN) { a.bf_list = S.bf.as<int32_t>(); a.bf_count = S.counters.as<int>(); }
            CU_TRY(c, launch_mark_keep(a, ps));
            CU_TRY(c, launch_count_keep_bits(a.keep_bits, B.staged_rows / 256, S.chunk_counts.as<int32_t>(), ps));
            c->stat_launches += need_cnt ? 3 : 2;
        } else {
            if (need_cnt) {
                CU_TRY(c, launch_filter_mark(a, ps));
                c->stat_launches++;
            }
            CU_TRY(c, launch_filter_count(a, S.chunk_counts.as<int32_t>(), ps));
            c->stat_launches++;
        }
        CU_TRY(c, launch_scan_offsets(S.chunk_counts.as<int32_t>(), B.staged_rows / 256, d_outp + base, np,
                                      o->min_match_count, S.chunk_excl.as<int64_t>(), S.pair_counts.as<int64_t>(),
                                      c->d_pair_offsets.as<int64_t>() + B.p0, c->d_dropped.as<uint8_t>() + B.p0, d_total, ps));
        if (sparse)
            CU_TRY(c, launch_compact_keep_bits(a, S.chunk_excl.as<int64_t>(), c->d_pair_offsets.as<int64_t>() + B.p0,
                                               c->d_dropped.as<uint8_t>() + B.p0, c->d_out.as<DMatch>(), c->out_capacity, d_overflow, ps));
        else
            CU_TRY(c, launch_compact(a, S.chunk_excl.as<int64_t>(), c->d_pair_offsets.as<int64_t>() + B.p0,
                                     c->d_dropped.as<uint8_t>() + B.p0, c->d_out.as<DMatch>(), c->out_capacity, d_overflow, ps));
        c->stat_launches += 2;
        if (c->profiling) { CU_TRY(c, cudaEventRecord(c->prof_ev[3 * bi + 2], ps)); c->prof_used = static_cast<int>(3 * (bi + 1)); }
        CU_TRY(c, cudaEventRecord(S.post_done, ps));
    }
    // join: everything after this on the compute stream (reorder, collect, homography, the next run) sees the finished lists
    if (nb > 0) CU_TRY(c, cudaStreamWaitEvent(s, c->slot[(nb - 1) & 1].post_done, 0));
    if (sched && sched->order && n_pairs > 0) {
        // schedule order -> input order (swap the buffers so that collect / device_view see input order)
        CU_TRY(c, c->d_order.ensure(static_cast<size_t>(n_pairs) * 8));
        CU_TRY(c, cudaMemcpyAsync(c->d_order.p, sched->order, static_cast<size_t>(n_pairs) * 8, cudaMemcpyHostToDevice, s));
        CU_TRY(c, c->d_cnt_tmp.ensure(static_cast<size_t>(n_pairs) * 8));
        CU_TRY(c, c->d_pair_offsets2.ensure(static_cast<size_t>(n_pairs) * 8));
        CU_TRY(c, c->d_dropped2.ensure(std::max<size_t>(16, static_cast<size_t>(n_pairs))));
        CU_TRY(c, c->d_out2.ensure(static_cast<size_t>(c->out_capacity) * sizeof(DMatch)));
        CU_TRY(c, launch_reorder(c->d_out.as<DMatch>(), c->d_pair_offsets.as<int64_t>(), d_total, c->d_order.as<int64_t>(),
                                 c->d_dropped.as<uint8_t>(), n_pairs, c->d_cnt_tmp.as<int64_t>(),
                                 c->d_pair_offsets2.as<int64_t>(), c->d_out2.as<DMatch>(), c->d_dropped2.as<uint8_t>(), s));
        c->stat_launches += 3;
        CU_TRY(c, cudaStreamSynchronize(s));          // sched->order is caller-owned host memory
        std::swap(c->d_out, c->d_out2);
        std::swap(c->d_pair_offsets, c->d_pair_offsets2);
        std::swap(c->d_dropped, c->d_dropped2);
    }
    if ((eng == Engine::TCV || eng == Engine::TCN) && total_q > 0) {
        CU_TRY(c, cudaMemcpyAsync(c->h_tune.p, c->d_scalars.as<uint8_t>() + 32, 16, cudaMemcpyDeviceToHost, s));
        CU_TRY(c, cudaEventRecord(c->tune_ev, s));
        c->tune_pending = true;
        c->tune_rows = total_q;
    }
    pairs = pairs_in;
    c->run.valid = true;
    c->run.n_pairs = n_pairs;
    c->run.total_query_rows = total_q;
    c->run.opts = *o;
    if (c->run.pairs.data() != pairs) c->run.pairs.assign(pairs, pairs + 2 * n_pairs);
    return SFM_OK;
}

int collect_impl(sfm_ctx* c, sfm_result** out) {
    NvtxRange nvtx_range("sfm:match_pairs_collect");
    if (!c->run.valid) return fail(c, SFM_ERR_STATE, "collect without a preceding enqueue");
    cudaStream_t s = c->stream;
    const int64_t n = c->run.n_pairs;
    for (int attempt = 0; attempt < 2; ++attempt) {
        int64_t* hs = c->h_scalars.as<int64_t>();
        CU_TRY(c, cudaMemcpyAsync(hs, c->d_scalars.p, 16, cudaMemcpyDeviceToHost, s));
        CU_TRY(c, cudaStreamSynchronize(s));
        const int64_t total = hs[0];
        const int overflow = reinterpret_cast<int*>(hs)[2];
        if (overflow) {
            if (attempt == 1) return fail(c, SFM_ERR_CAPACITY, "output capacity overflow after retry");
            c->out_capacity = std::max<int64_t>(c->run.total_query_rows, 1);
            std::vector<int32_t> keep;
            keep.swap(c->run.pairs);
            sfm_opts o = c->run.opts;
            int rc = enqueue_impl(c, keep.data(), n, &o);
            c->run.pairs.swap(keep);
            if (rc != SFM_OK) return rc;
            continue;
        }
        sfm_result* r;
        if (!c->result_pool.empty()) { r = c->result_pool.back(); c->result_pool.pop_back(); }
        else r = new sfm_result();
        r->owner = c;
        r->n_pairs = n;
        auto fill = [&]() -> int {
            CU_TRY(c, r->offsets.ensure(static_cast<size_t>(n + 1) * 8));
            CU_TRY(c, r->dropped.ensure(std::max<size_t>(16, static_cast<size_t>(n))));
            CU_TRY(c, r->matches.ensure(std::max<size_t>(16, static_cast<size_t>(total) * sizeof(DMatch))));
            if (n > 0) {
                CU_TRY(c, cudaMemcpyAsync(r->offsets.p, c->d_pair_offsets.p, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost, s));
                CU_TRY(c, cudaMemcpyAsync(r->dropped.p, c->d_dropped.p, static_cast<size_t>(n), cudaMemcpyDeviceToHost, s));
            }
            if (total > 0)
                CU_TRY(c, cudaMemcpyAsync(r->matches.p, c->d_out.p, static_cast<size_t>(total) * sizeof(DMatch), cudaMemcpyDeviceToHost, s));
            CU_TRY(c, cudaStreamSynchronize(s));
            return SFM_OK;
        };
        const int rc_fill = fill();
        if (rc_fill != SFM_OK) { c->result_pool.push_back(r); return rc_fill; }      // keep the result object for the next run
        r->offsets.as<int64_t>()[n] = total;
        c->stat_d2h += 16 + n * 9 + total * static_cast<int64_t>(sizeof(DMatch));
        *out = r;
        return SFM_OK;
    }
    return fail(c, SFM_ERR_CAPACITY, "unreachable");
}

// ------------------------------------------------------------------------------------------------ pipelined host path
// sfm_match_pairs_from_host: upload + match + collect as one call.  With page-locked 128-column descriptors the images
// are uploaded in groups on a copy stream (H2D, pack to u8, norms/keys/digits per group, one event per group) while the
// compute stream already matches the pairs whose two images are resident; pairs are scheduled by availability and the
// lists are brought back to input order at the end.  Bank properties (integer-valued, norm range) are assumed and
// verified after the fact; if the assumption fails the call falls back to the sequential path.
static int from_host_body(sfm_ctx* c, int n_images, const void* const* rows, const int32_t* n_rows, int cols,
                          const size_t* step_bytes, int depth, const int32_t* pairs, int64_t n_pairs, const sfm_opts* o,
                          sfm_result** out);

int from_host_impl(sfm_ctx* c, int n_images, const void* const* rows, const int32_t* n_rows, int cols,
                   const size_t* step_bytes, int depth, const int32_t* pairs, int64_t n_pairs, const sfm_opts* o,
                   sfm_result** out) {
    const int rc = from_host_body(c, n_images, rows, n_rows, cols, step_bytes, depth, pairs, n_pairs, o, out);
    if (rc != SFM_OK) {
        // the pipelined path sets the bank's properties before the data has been verified: after a failure nothing may
        // look resident (a later sfm_match_pairs would match against unverified or partly uploaded rows)
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamSynchronize(c->stream);
        c->bank.n_images = 0; c->bank.have_tmap = false; c->bank.n_rows.clear();
        c->run.valid = false;
    }
    return rc;
}

static int from_host_body(sfm_ctx* c, int n_images, const void* const* rows, const int32_t* n_rows, int cols,
                          const size_t* step_bytes, int depth, const int32_t* pairs, int64_t n_pairs, const sfm_opts* o,
                          sfm_result** out) {
    auto sequential = [&]() -> int {
        int rc = bank_upload_host(c, c->bank, n_images, rows, n_rows, cols, step_bytes, depth);
        if (rc != SFM_OK) return rc;
        rc = enqueue_impl(c, pairs, n_pairs, o);
        if (rc != SFM_OK) return rc;
        return collect_impl(c, out);
    };
    bool pipelined = cols == 128 && (depth == SFM_CV_32F || depth == SFM_CV_8U) && o->norm == SFM_NORM_L2 && o->k == 2 &&
                     !o->cross_check && (o->engine == SFM_ENGINE_AUTO || o->engine == SFM_ENGINE_TENSOR) &&
                     n_images >= 16 && n_pairs > 0 && pairs && rows && n_rows;
    const size_t esz = depth == SFM_CV_32F ? 4 : 1;
    const size_t row_bytes = static_cast<size_t>(cols) * esz;
    for (int i = 0; pipelined && i < n_images; ++i) {
        if (n_rows[i] < 0 || n_rows[i] > 32768) { pipelined = false; break; }
        if (n_rows[i] == 0) continue;
        cudaPointerAttributes attr;
        if (!rows[i] || cudaPointerGetAttributes(&attr, rows[i]) != cudaSuccess || attr.type != cudaMemoryTypeHost) {
            cudaGetLastError();
            pipelined = false;
        }
        if (step_bytes && step_bytes[i] < row_bytes) pipelined = false;
    }
    for (int64_t p = 0; pipelined && p < n_pairs; ++p) {
        const int l = pairs[2 * p], r = pairs[2 * p + 1];
        if (l < 0 || r < 0 || l >= n_images || r >= n_images) pipelined = false;       // let the sequential path report it
    }
    if (!pipelined) return sequential();

    Bank& b = c->bank;
    int rc = bank_layout(c, b, n_images, n_rows, cols, depth);
    if (rc != SFM_OK) return rc;
    if (b.padded_rows == 0) return sequential();
    cudaStream_t cs = c->copy_stream, s = c->stream;
    CU_TRY(c, cudaStreamSynchronize(s));                  // earlier work may still read the buffers we are about to refill
    // ---- buffers + valid table
    const int64_t nblk = b.padded_rows / kRowAlign;
    CU_TRY(c, cudaEventSynchronize(c->valid_ev));
    CU_TRY(c, c->h_valid.ensure(static_cast<size_t>(nblk) * 4));
    int32_t* valid = c->h_valid.as<int32_t>();
    for (int i = 0; i < n_images; ++i) {
        const int64_t b0 = b.row0[i] / kRowAlign, nbk = pad_rows(b.n_rows[i]) / kRowAlign;
        for (int64_t k = 0; k < nbk; ++k)
            valid[b0 + k] = static_cast<int32_t>(std::min<int64_t>(kRowAlign, std::max<int64_t>(0, b.n_rows[i] - k * kRowAlign)));
    }
    CU_TRY(c, b.d_valid.ensure(static_cast<size_t>(nblk) * 4));
    CU_TRY(c, b.d_u8.ensure(static_cast<size_t>(b.padded_rows) * 128));
    if (depth == SFM_CV_32F) CU_TRY(c, b.d_f32.ensure(static_cast<size_t>(b.padded_rows) * 512));
    CU_TRY(c, b.d_norm2.ensure(b.padded_rows * 4));
    CU_TRY(c, b.d_ckey.ensure(b.padded_rows * 4));
    CU_TRY(c, b.d_ext.ensure(static_cast<size_t>(b.padded_rows) * kExtBytes));
    CU_TRY(c, cudaMemcpyAsync(b.d_valid.p, valid, static_cast<size_t>(nblk) * 4, cudaMemcpyHostToDevice, cs));
    CU_TRY(c, cudaEventRecord(c->valid_ev, cs));
    int* flags = reinterpret_cast<int*>(c->d_scalars.as<uint8_t>() + 16);
    CU_TRY(c, cudaMemsetAsync(flags, 0, 8, cs));
    CU_TRY(c, cudaMemsetAsync(flags + 2, 0x7f, 4, cs));
    const size_t nblk_b = static_cast<size_t>(nblk) * 4;
    CU_TRY(c, b.d_blkmin.ensure(nblk_b));
    CU_TRY(c, b.d_blkmax.ensure(nblk_b));
    CU_TRY(c, cudaMemsetAsync(b.d_blkmin.p, 0x7f, nblk_b, cs));
    CU_TRY(c, cudaMemsetAsync(b.d_blkmax.p, 0, nblk_b, cs));
    // optimistic bank properties (verified below): integer-valued SIFT-sized rows.  The norm spread is unknown until the
    // data has arrived: assume that of the previous bank of this context (the first call gets the kernel with the norm
    // K-step); a wrong guess only costs time in the refine pass, never exactness
    b.u8_valued = true; b.have_f32 = false; b.ext_ok = true; b.nb_min = c->prev_nb_min; b.nb_max = c->prev_nb_max;
    rc = make_tmaps(c, b);
    if (rc != SFM_OK) return rc;
    // ---- image groups of roughly equal size
    const int n_groups = std::min(8, n_images / 2);
    std::vector<int> group_of(n_images);
    std::vector<int> group_end(n_groups);
    {
        int g = 0;
        for (int i = 0; i < n_images; ++i) {
            group_of[i] = g;
            group_end[g] = i + 1;
            if (b.row0[i + 1] * n_groups >= b.padded_rows * (g + 1) && g + 1 < n_groups && i + 1 < n_images) ++g;
        }
        for (int k = g + 1; k < n_groups; ++k) group_end[k] = n_images;
    }
    while (static_cast<int>(c->group_ev.size()) < n_groups) {
        cudaEvent_t ev;
        CU_TRY(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        c->group_ev.push_back(ev);
    }
    const int32_t* d_valid = b.d_valid.as<int32_t>();
    int first = 0;
    for (int g = 0; g < n_groups; ++g) {
        const int last = group_end[g];
        const int64_t r0 = b.row0[first], r1 = b.row0[last];
        uint8_t* dst = static_cast<uint8_t*>(depth == SFM_CV_32F ? b.d_f32.p : b.d_u8.p);
        for (int i = first; i < last; ++i) {
            if (n_rows[i] == 0) continue;
            const size_t step = step_bytes ? step_bytes[i] : row_bytes;
            uint8_t* d_img = dst + static_cast<size_t>(b.row0[i]) * row_bytes;
            if (step == row_bytes)
                CU_TRY(c, cudaMemcpyAsync(d_img, rows[i], static_cast<size_t>(n_rows[i]) * row_bytes, cudaMemcpyHostToDevice, cs));
            else
                CU_TRY(c, cudaMemcpy2DAsync(d_img, row_bytes, rows[i], step, row_bytes, static_cast<size_t>(n_rows[i]),
                                            cudaMemcpyHostToDevice, cs));
            c->stat_h2d += static_cast<int64_t>(n_rows[i]) * static_cast<int64_t>(row_bytes);
        }
        if (r1 > r0) {
            if (depth == SFM_CV_32F)
                CU_TRY(c, launch_pack_f32_to_u8(b.d_f32.as<float>() + r0 * 128, 128, static_cast<int>(r1 - r0), 128,
                                                d_valid + r0 / kRowAlign, b.d_u8.as<uint8_t>() + r0 * 128, flags, cs));
            else
                CU_TRY(c, launch_zero_padding(b.d_u8.as<uint8_t>() + r0 * 128, 128, r1 - r0, d_valid + r0 / kRowAlign, cs));
            CU_TRY(c, launch_norms_ckeys(b.d_u8.as<uint8_t>() + r0 * 128, r1 - r0, d_valid + r0 / kRowAlign,
                                         b.d_norm2.as<int32_t>() + r0, b.d_ckey.as<int32_t>() + r0,
                                         b.d_ext.as<int8_t>() + r0 * kExtBytes, flags + 1,
                                         b.d_blkmin.as<int32_t>() + r0 / kRowAlign, b.d_blkmax.as<int32_t>() + r0 / kRowAlign, cs));
            c->stat_launches += 2;
        }
        CU_TRY(c, cudaEventRecord(c->group_ev[g], cs));
        first = last;
    }
    int* h = c->h_scalars.as<int>() + 8;
    CU_TRY(c, cudaMemcpyAsync(h, flags, 12, cudaMemcpyDeviceToHost, cs));
    // ---- schedule pairs by the group that completes them
    std::vector<int64_t> order(n_pairs);
    std::vector<int> avail_in(n_pairs), avail(n_pairs);
    for (int64_t p = 0; p < n_pairs; ++p) {
        order[p] = p;
        avail_in[p] = std::max(group_of[pairs[2 * p]], group_of[pairs[2 * p + 1]]);
    }
    std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return avail_in[x] < avail_in[y]; });
    for (int64_t k = 0; k < n_pairs; ++k) avail[k] = avail_in[order[k]];
    Schedule sc;
    sc.order = order.data(); sc.avail = avail.data(); sc.events = c->group_ev.data();
    rc = enqueue_impl(c, pairs, n_pairs, o, &sc);
    CU_TRY(c, cudaStreamSynchronize(cs));
    if (rc != SFM_OK) return rc;
    b.nb_max = h[1]; b.nb_min = std::min(h[2], h[1]);
    c->prev_nb_min = b.nb_min; c->prev_nb_max = b.nb_max;
    if (h[0] != 0 || h[1] > kExtMaxNorm2) {
        // not integer-valued, or norms beyond the digit range: the optimistic run is void -> sequential path
        CU_TRY(c, cudaStreamSynchronize(s));
        c->run.valid = false;
        return sequential();
    }
    return collect_impl(c, out);
}

}  // namespace sfmhost

// ==================================================================================================== C ABI
extern "C" {

void sfm_opts_default(sfm_opts* o, int32_t norm) {
    o->norm = norm; o->k = 2; o->ratio = 0.7; o->cross_check = 0; o->distinct = 0; o->min_match_count = 0;
    o->engine = SFM_ENGINE_AUTO;
}

int sfm_ctx_create(sfm_ctx** out, int device) {
    if (!out) return fail(nullptr, SFM_ERR_INVALID, "null out pointer");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, SFM_ERR_CUDA, std::string("no CUDA device usable (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, SFM_ERR_INVALID, "device index out of range");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(nullptr, SFM_ERR_CUDA, cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, SFM_ERR_CUDA, "device is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                                               "; this library carries sm_100a code only");
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, SFM_ERR_CUDA, cudaGetErrorString(e));
    sfm_ctx* c = new sfm_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    auto bail = [&](const std::string& m) { g_create_error = m; sfm_ctx_destroy(c); return SFM_ERR_CUDA; };
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(cudaGetErrorString(e));
    if ((e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(cudaGetErrorString(e));
    if ((e = cudaStreamCreateWithFlags(&c->post_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(cudaGetErrorString(e));
    for (int k = 0; k < 2; ++k) {
        if ((e = cudaEventCreateWithFlags(&c->slot[k].knn_done, cudaEventDisableTiming)) != cudaSuccess) return bail(cudaGetErrorString(e));
        if ((e = cudaEventCreateWithFlags(&c->slot[k].post_done, cudaEventDisableTiming)) != cudaSuccess) return bail(cudaGetErrorString(e));
    }
    for (int k = 0; k < 2; ++k)
        if ((e = cudaEventCreateWithFlags(&c->stage_ev[k], cudaEventDisableTiming)) != cudaSuccess) return bail(cudaGetErrorString(e));
    if ((e = cudaEventCreateWithFlags(&c->meta_ev, cudaEventDisableTiming)) != cudaSuccess) return bail(cudaGetErrorString(e));
    if ((e = cudaEventCreateWithFlags(&c->valid_ev, cudaEventDisableTiming)) != cudaSuccess) return bail(cudaGetErrorString(e));
    if ((e = cudaEventCreateWithFlags(&c->tune_ev, cudaEventDisableTiming)) != cudaSuccess) return bail(cudaGetErrorString(e));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !fn) return bail("cuTensorMapEncodeTiled not available from the driver");
    c->encode = reinterpret_cast<EncodeTiledFn>(fn);
    if ((e = c->d_scalars.ensure(64)) != cudaSuccess) return bail(cudaGetErrorString(e));
    if ((e = c->h_scalars.ensure(64)) != cudaSuccess) return bail(cudaGetErrorString(e));
    if ((e = c->h_tune.ensure(64)) != cudaSuccess) return bail(cudaGetErrorString(e));
    size_t mb = 512;
    if (const char* env = std::getenv("SFM_STAGING_MB")) { long v = std::atol(env); if (v > 0) mb = static_cast<size_t>(v); }
    c->staging_budget_rows = (mb << 20) / sizeof(Top2);
    if (const char* env = std::getenv("SFM_TCV_INKERNEL_REFINE")) c->tcv_inkernel_refine = std::atoi(env) != 0;
    if (const char* env = std::getenv("SFM_TCV_BACKPRESSURE")) { const int t = std::atoi(env); if (t >= 0 && t <= 2) c->tcv_backpressure = t; }
    if (const char* env = std::getenv("SFM_TCV_DIVERT_TEST")) c->tcv_divert_test = std::atoi(env) != 0;
    if (const char* env = std::getenv("SFM_MIN_BATCHES")) { const int t = std::atoi(env); if (t >= 1 && t <= 64) c->min_batches = t; }
    if (const char* env = std::getenv("SFM_TCV_SPREAD_DIV")) { const int t = std::atoi(env); if (t >= 1) c->tcv_spread_div = t; }
    if (const char* env = std::getenv("SFM_TCV_NORMLESS")) { const int t = std::atoi(env); if (t >= 0 && t <= 2) c->tcv_normless = t; }
    if (const char* env = std::getenv("SFM_TCV_CHUNK")) { const int t = std::atoi(env); if (t == 32 || t == 64 || t == 128) c->tcv_chunk = t; }
    if (const char* env = std::getenv("SFM_TCV_ISSUERS")) { const int t = std::atoi(env); if (t == 1 || t == 2) c->tcv_issuers = t; }
    if (const char* env = std::getenv("SFM_TCV_LAYOUT")) { const int g = std::atoi(env); if (g == 12 || g == 14 || g == 21 || g == 22) c->tcv_layout = g; }
    *out = c;
    return SFM_OK;
}

void sfm_ctx_destroy(sfm_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->post_stream) cudaStreamSynchronize(c->post_stream);
    dist_state_destroy(c);
    c->slot[0].release(); c->slot[1].release();
    if (c->post_stream) cudaStreamDestroy(c->post_stream);
    c->bank.release(); c->scratch.release();
    DevBuf* bufs[] = {&c->d_pairs, &c->d_rev_pairs, &c->d_unit_prefix, &c->d_rev_unit_prefix, &c->d_out_prefix, &c->d_t_prefix,
                      &c->d_pair_offsets, &c->d_dropped, &c->d_scalars, &c->d_out, &c->d_knn,
                      &c->d_out2, &c->d_pair_offsets2, &c->d_dropped2, &c->d_order, &c->d_cnt_tmp, &c->d_hom};
    for (DevBuf* b : bufs) b->release();
    c->feat_kp.release(); c->feat_desc.release();
    sift_workspace_destroy(c->sift);
    orb_workspace_destroy(c->orb);
    c->h_meta.release(); c->h_stage[0].release(); c->h_stage[1].release(); c->h_scalars.release(); c->h_knn.release();
    for (int k = 0; k < 2; ++k) if (c->stage_ev[k]) cudaEventDestroy(c->stage_ev[k]);
    if (c->meta_ev) cudaEventDestroy(c->meta_ev);
    if (c->valid_ev) cudaEventDestroy(c->valid_ev);
    if (c->tune_ev) cudaEventDestroy(c->tune_ev);
    c->h_tune.release();
    for (sfm_result* r : c->result_pool) { r->offsets.release(); r->matches.release(); r->dropped.release(); delete r; }
    c->h_valid.release();
    for (cudaEvent_t ev : c->prof_ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : c->group_ev) cudaEventDestroy(ev);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* sfm_last_error(const sfm_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }
int sfm_device_sm_count(const sfm_ctx* c) { return c ? c->sm_count : 0; }
void* sfm_ctx_stream(sfm_ctx* c) { return c ? static_cast<void*>(c->stream) : nullptr; }

int sfm_bank_upload(sfm_ctx* c, int n_images, const void* const* rows, const int32_t* n_rows, int cols,
                    const size_t* step_bytes, int cv_depth) {
    if (!c) return SFM_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (n_images > 0 && (!rows || !n_rows)) return fail(c, SFM_ERR_INVALID, "bank: null arrays");
    CU_TRY(c, cudaSetDevice(c->device));
    c->run.valid = false;
    return bank_upload_host(c, c->bank, n_images, rows, n_rows, cols, step_bytes, cv_depth);
}

int sfm_bank_upload_device(sfm_ctx* c, int n_images, const void* dev_rows, const int64_t* row_offset, const int32_t* n_rows,
                           int cols, int cv_depth) {
    if (!c) return SFM_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (n_images > 0 && (!dev_rows || !n_rows || !row_offset)) return fail(c, SFM_ERR_INVALID, "bank: null arrays");
    CU_TRY(c, cudaSetDevice(c->device));
    c->run.valid = false;
    Bank& b = c->bank;
    int rc = bank_layout(c, b, n_images, n_rows, cols, cv_depth);
    if (rc != SFM_OK) return rc;
    const size_t esz = cv_depth == SFM_CV_32F ? 4 : 1;
    const size_t row_bytes = static_cast<size_t>(cols) * esz;
    DevBuf& dst = cv_depth == SFM_CV_32F ? b.d_f32 : b.d_u8;
    CU_TRY(c, dst.ensure(std::max<size_t>(16, static_cast<size_t>(b.padded_rows) * row_bytes)));
    bool same_layout = true;          // source already laid out like the padded bank: one copy instead of n_images
    for (int i = 0; i < n_images && same_layout; ++i) same_layout = row_offset[i] == b.row0[i];
    if (same_layout && n_images > 0 && b.padded_rows > 0) {
        const int64_t last = b.row0[n_images - 1] + n_rows[n_images - 1];
        CU_TRY(c, cudaMemcpyAsync(dst.p, dev_rows, static_cast<size_t>(last) * row_bytes, cudaMemcpyDeviceToDevice, c->stream));
    }
    for (int i = 0; i < n_images && !same_layout; ++i) {
        if (n_rows[i] == 0) continue;
        CU_TRY(c, cudaMemcpyAsync(static_cast<uint8_t*>(dst.p) + static_cast<size_t>(b.row0[i]) * row_bytes,
                                  static_cast<const uint8_t*>(dev_rows) + static_cast<size_t>(row_offset[i]) * row_bytes,
                                  static_cast<size_t>(n_rows[i]) * row_bytes, cudaMemcpyDeviceToDevice, c->stream));
    }
    return bank_finish(c, b);
}

int sfm_bank_device_ptr(const sfm_ctx* c, int image, const void** dev_rows, int32_t* n_rows) {
    if (!c || !dev_rows) return SFM_ERR_INVALID;
    const Bank& b = c->bank;
    if (image < 0 || image >= b.n_images || !b.u8_valued) return SFM_ERR_STATE;
    *dev_rows = b.d_u8.as<uint8_t>() + static_cast<size_t>(b.row0[image]) * b.cols;
    if (n_rows) *n_rows = b.n_rows[image];
    return SFM_OK;
}

int sfm_bank_info(const sfm_ctx* c, int* n_images, int* cols, int* is_u8_valued) {
    if (!c) return SFM_ERR_INVALID;
    if (n_images) *n_images = c->bank.n_images;
    if (cols) *cols = c->bank.cols;
    if (is_u8_valued) *is_u8_valued = c->bank.u8_valued ? 1 : 0;
    return SFM_OK;
}

int sfm_match_pairs_enqueue(sfm_ctx* c, const int32_t* pairs, int64_t n_pairs, const sfm_opts* opts) {
    if (!c || !opts) return SFM_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    CU_TRY(c, cudaSetDevice(c->device));
    c->run.valid = false;
    return enqueue_impl(c, pairs, n_pairs, opts);
}

int sfm_match_pairs_device_view(sfm_ctx* c, const void** d_matches, const void** d_pair_offsets, const void** d_dropped,
                                int64_t* n_pairs, int64_t* total_matches) {
    if (!c) return SFM_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->run.valid) return fail(c, SFM_ERR_STATE, "device view without a preceding enqueue");
    CU_TRY(c, cudaSetDevice(c->device));
    int64_t* hs = c->h_scalars.as<int64_t>();
    CU_TRY(c, cudaMemcpyAsync(hs, c->d_scalars.p, 16, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->stat_d2h += 16;
    if (reinterpret_cast<int*>(hs)[2]) return fail(c, SFM_ERR_CAPACITY, "output capacity overflow: use sfm_match_pairs_collect (it retries)");
    if (d_matches) *d_matches = c->d_out.p;
    if (d_pair_offsets) *d_pair_offsets = c->d_pair_offsets.p;
    if (d_dropped) *d_dropped = c->d_dropped.p;
    if (n_pairs) *n_pairs = c->run.n_pairs;
    if (total_matches) *total_matches = hs[0];
    return SFM_OK;
}

int sfm_match_pairs_collect(sfm_ctx* c, sfm_result** out) {
    if (!c || !out) return SFM_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    CU_TRY(c, cudaSetDevice(c->device));
    return collect_impl(c, out);
}

int sfm_match_pairs(sfm_ctx* c, const int32_t* pairs, int64_t n_pairs, const sfm_opts* opts, sfm_result** out) {
    int rc = sfm_match_pairs_enqueue(c, pairs, n_pairs, opts);
    if (rc != SFM_OK) return rc;
    return sfm_match_pairs_collect(c, out);
}

int sfm_match_pairs_from_host(sfm_ctx* c, int n_images, const void* const* rows, const int32_t* n_rows, int cols,
                              const size_t* step_bytes, int cv_depth, const int32_t* pairs, int64_t n_pairs,
                              const sfm_opts* opts, sfm_result** out) {
    if (!c || !opts || !out) return SFM_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (n_images > 0 && (!rows || !n_rows)) return fail(c, SFM_ERR_INVALID, "bank: null arrays");
    CU_TRY(c, cudaSetDevice(c->device));
    c->run.valid = false;
    return from_host_impl(c, n_images, rows, n_rows, cols, step_bytes, cv_depth, pairs, n_pairs, opts, out);
}

int64_t sfm_result_n_pairs(const sfm_result* r) { return r ? r->n_pairs : 0; }
const int64_t* sfm_result_offsets(const sfm_result* r) { return r ? r->offsets.as<int64_t>() : nullptr; }
const sfm_dmatch* sfm_result_matches(const sfm_result* r) { return r ? r->matches.as<sfm_dmatch>() : nullptr; }
const uint8_t* sfm_result_dropped(const sfm_result* r) { return r ? r->dropped.as<uint8_t>() : nullptr; }
void sfm_result_free(sfm_result* r) {
    if (!r) return;
    sfm_ctx* c = r->owner;
    if (c) {
        std::lock_guard<std::mutex> lk(c->mu);
        if (c->result_pool.size() < 4) { c->result_pool.push_back(r); return; }
    }
    r->offsets.release(); r->matches.release(); r->dropped.release();
    delete r;
}

int sfm_last_stats(const sfm_ctx* c, int64_t* launches, int64_t* h2d, int64_t* d2h) {
    if (!c) return SFM_ERR_INVALID;
    if (launches) *launches = c->stat_launches;
    if (h2d) *h2d = c->stat_h2d;
    if (d2h) *d2h = c->stat_d2h;
    return SFM_OK;
}

int sfm_last_float_stats(sfm_ctx* c, int64_t* rows_reranked, int64_t* rows_brute_forced) {
    if (!c) return SFM_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    CU_TRY(c, cudaSetDevice(c->device));
    int64_t h[2];
    CU_TRY(c, cudaMemcpyAsync(h, c->d_scalars.as<uint8_t>() + 32, 16, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    if (rows_reranked) *rows_reranked = h[0];
    if (rows_brute_forced) *rows_brute_forced = h[1];
    return SFM_OK;
}

// ---- SfM::calculateHomography (SfM.cpp:599-637) on the device-resident match lists
int sfm_keypoints_upload(sfm_ctx* c, int n_images, const void* const* pts, const int32_t* n_rows, const size_t* step_bytes) {
    if (!c) return SFM_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    Bank& b = c->bank;
    if (n_images != b.n_images) return fail(c, SFM_ERR_STATE, "keypoints: image count differs from the descriptor bank");
    if (n_images > 0 && (!pts || !n_rows)) return fail(c, SFM_ERR_INVALID, "keypoints: null arrays");
    for (int i = 0; i < n_images; ++i)
        if (n_rows[i] != b.n_rows[i]) return fail(c, SFM_ERR_INVALID, "keypoints: row count differs from the descriptor bank");
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, b.d_kp.ensure(std::max<size_t>(16, static_cast<size_t>(b.padded_rows) * 8)));
    for (int i = 0; i < n_images; ++i) {
        if (n_rows[i] == 0) continue;
        if (!pts[i]) return fail(c, SFM_ERR_INVALID, "keypoints: null image pointer");
        const size_t step = step_bytes && step_bytes[i] ? step_bytes[i] : 8;       // sizeof(cv::KeyPoint) = 28 for a KeyPoint vector
        if (step < 8) return fail(c, SFM_ERR_INVALID, "keypoints: step < 8 bytes");
        CU_TRY(c, cudaMemcpy2DAsync(b.d_kp.as<uint8_t>() + static_cast<size_t>(b.row0[i]) * 8, 8, pts[i], step, 8,
                                    static_cast<size_t>(n_rows[i]), cudaMemcpyHostToDevice, c->stream));
        c->stat_h2d += static_cast<int64_t>(n_rows[i]) * 8;
    }
    CU_TRY(c, cudaStreamSynchronize(c->stream));       // the caller's buffers may go away
    b.have_kp = true;
    return SFM_OK;
}

void sfm_homography_opts_default(sfm_homography_opts* o) {
    if (!o) return;
    o->max_iters = 2000;       // cv::findHomography defaults (calib3d.hpp), as SfM.cpp:622-625 leaves them
    o->confidence = 0.995;
    o->refine = 1;
    o->seed = 0;
}

int sfm_homography_inlier_ratios(sfm_ctx* c, const double* thresholds, int64_t n_thresholds, const sfm_homography_opts* opts,
                                 double* ratios, int32_t* inliers, int32_t* ransac_inliers, int32_t* best_hypothesis) {
    if (!c) return SFM_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    sfm_homography_opts o;
    sfm_homography_opts_default(&o);
    if (opts) o = *opts;
    NvtxRange nvtx_range("sfm:homography_inlier_ratios");
    if (!c->run.valid) return fail(c, SFM_ERR_STATE, "homography without a preceding match_pairs run");
    if (!c->bank.have_kp) return fail(c, SFM_ERR_STATE, "homography before sfm_keypoints_upload");
    const int64_t n = c->run.n_pairs;
    if (!thresholds || (n_thresholds != 1 && n_thresholds != n)) return fail(c, SFM_ERR_INVALID, "homography: one threshold, or one per pair");
    if (o.max_iters <= 0 || o.max_iters > 16384) return fail(c, SFM_ERR_INVALID, "homography: max_iters must be in 1..16384");
    if (!(o.confidence > 0.0)) return fail(c, SFM_ERR_INVALID, "homography: confidence must be > 0");
    for (int64_t i = 0; i < n_thresholds; ++i)
        if (!(thresholds[i] > 0.0)) return fail(c, SFM_ERR_INVALID, "homography: threshold must be > 0 pixels");
    if (n == 0) return SFM_OK;
    if (!ratios) return fail(c, SFM_ERR_INVALID, "homography: null output");
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const Bank& b = c->bank;
    // d_hom: row0[2n] int64 | thresholds[nt] double | inliers[n] | ransac inliers[n] | best hypothesis[n] (int32)
    const size_t off_thr = static_cast<size_t>(2 * n) * 8, off_out = off_thr + static_cast<size_t>(n_thresholds) * 8;
    const size_t bytes = off_out + static_cast<size_t>(3 * n) * 4;
    CU_TRY(c, c->d_hom.ensure(bytes));
    std::vector<uint8_t> h(off_out);
    int64_t* h_row0 = reinterpret_cast<int64_t*>(h.data());
    for (int64_t p = 0; p < n; ++p) {
        h_row0[2 * p] = b.row0[c->run.pairs[2 * p]];
        h_row0[2 * p + 1] = b.row0[c->run.pairs[2 * p + 1]];
    }
    std::memcpy(h.data() + off_thr, thresholds, static_cast<size_t>(n_thresholds) * 8);
    CU_TRY(c, cudaMemcpyAsync(c->d_hom.p, h.data(), off_out, cudaMemcpyHostToDevice, s));
    c->stat_h2d += static_cast<int64_t>(off_out);
    HomographyArgs a;
    a.keypoints = b.d_kp.as<float2>();
    a.row0 = c->d_hom.as<int64_t>();
    a.matches = c->d_out.as<DMatch>();
    a.pair_offsets = c->d_pair_offsets.as<int64_t>();
    a.total = c->d_scalars.as<int64_t>();
    a.dropped = c->d_dropped.as<uint8_t>();
    a.n_pairs = static_cast<int>(n);
    a.thresholds = reinterpret_cast<const double*>(c->d_hom.as<uint8_t>() + off_thr);
    a.n_thresholds = static_cast<int>(n_thresholds);
    a.max_iters = o.max_iters;
    a.confidence = o.confidence;
    a.refine = o.refine ? 1 : 0;
    a.seed = o.seed;
    a.inliers = reinterpret_cast<int32_t*>(c->d_hom.as<uint8_t>() + off_out);
    a.ransac_inliers = a.inliers + n;
    a.best_hyp = a.inliers + 2 * n;
    CU_TRY(c, launch_homography_ransac(a, s));
    c->stat_launches++;
    std::vector<int32_t> h_out(static_cast<size_t>(3 * n));
    std::vector<int64_t> h_off(static_cast<size_t>(n + 1));
    CU_TRY(c, cudaMemcpyAsync(h_out.data(), a.inliers, static_cast<size_t>(3 * n) * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(c, cudaMemcpyAsync(h_off.data(), c->d_pair_offsets.p, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost, s));
    CU_TRY(c, cudaMemcpyAsync(&h_off[n], c->d_scalars.p, 8, cudaMemcpyDeviceToHost, s));
    CU_TRY(c, cudaStreamSynchronize(s));
    c->stat_d2h += static_cast<int64_t>(3 * n) * 4 + (n + 1) * 8;
    for (int64_t p = 0; p < n; ++p) {
        const int32_t k = h_out[p];
        const int64_t m = h_off[p + 1] - h_off[p];
        // inlierCount / matches.size(); pairs without a homography keep ShotMatches' initial -1 (Scene.h:56)
        ratios[p] = k < 0 ? -1.0 : static_cast<double>(k) / static_cast<double>(m);
        if (inliers) inliers[p] = k;
        if (ransac_inliers) ransac_inliers[p] = h_out[n + p];
        if (best_hypothesis) best_hypothesis[p] = h_out[2 * n + p];
    }
    return SFM_OK;
}

// ---- SfM::extractFeatures (SfM.cpp:577-597) with cv::SIFT (PhotogrammetrieCli.cpp:342-357) on the device
static_assert(sizeof(sfm_keypoint) == 24, "sfm_keypoint layout (= sift::Keypoint of csrc/sift_core.cuh)");
static_assert(sizeof(sfm_sift_opts) == 40, "sfm_sift_opts layout");

}  // extern "C"

// one extracted image joins the context's device-resident feature set (all images of a set share the descriptor width)
static int append_features(sfm_ctx* c, const void* d_kp, const uint8_t* d_desc, int n, int desc_bytes) {
    if (c->feat_cols != 0 && c->feat_cols != desc_bytes && c->feat_off.size() > 1)
        return fail(c, SFM_ERR_STATE, "features: SIFT and ORB images cannot share one feature set (sfm_features_clear first)");
    c->feat_cols = desc_bytes;
    const int64_t used = c->feat_off.back();
    const size_t db = static_cast<size_t>(desc_bytes);
    CU_TRY(c, c->feat_kp.ensure_keep(static_cast<size_t>(used + n) * sizeof(sfm_keypoint) + 16, static_cast<size_t>(used) * sizeof(sfm_keypoint), c->stream));
    CU_TRY(c, c->feat_desc.ensure_keep(static_cast<size_t>(used + n) * db + 16, static_cast<size_t>(used) * db, c->stream));
    if (n > 0) {
        CU_TRY(c, cudaMemcpyAsync(c->feat_kp.as<sfm_keypoint>() + used, d_kp, static_cast<size_t>(n) * sizeof(sfm_keypoint), cudaMemcpyDeviceToDevice, c->stream));
        CU_TRY(c, cudaMemcpyAsync(c->feat_desc.as<uint8_t>() + used * db, d_desc, static_cast<size_t>(n) * db, cudaMemcpyDeviceToDevice, c->stream));
    }
    c->feat_off.push_back(used + n);
    return SFM_OK;
}

extern "C" {

void sfm_sift_opts_default(sfm_sift_opts* o) {
    if (!o) return;
    o->n_octave_layers = 3;          // cv::SIFT::create defaults (features2d.hpp); the reference CLI passes (0, 3, 0.09)
    o->max_keypoints = 0;            // 0: the matcher's per-image limit
    o->n_features = 0;               // cv::SIFT::create(nfeatures = 0): keep all; the reference passes its feature-limit
    o->reserved = 0;
    o->contrast_threshold = 0.04;
    o->edge_threshold = 10.0;
    o->sigma = 1.6;
}

int sfm_features_clear(sfm_ctx* c) {
    if (!c) return SFM_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    c->feat_off.assign(1, 0);
    c->feat_cols = 0;
    return SFM_OK;
}

int sfm_features_extract_sift(sfm_ctx* c, const uint8_t* gray, int rows, int cols, size_t step_bytes, const sfm_sift_opts* opts,
                              int32_t* n_keypoints) {
    if (!c) return SFM_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    NvtxRange nvtx_range("sfm:features_extract_sift");
    sfm_sift_opts o;
    sfm_sift_opts_default(&o);
    if (opts) o = *opts;
    if (!gray || rows <= 0 || cols <= 0) return fail(c, SFM_ERR_INVALID, "extract: empty image");
    if (step_bytes == 0) step_bytes = static_cast<size_t>(cols);
    if (step_bytes < static_cast<size_t>(cols)) return fail(c, SFM_ERR_INVALID, "extract: step smaller than a row");
    if (rows > 32768 || cols > 32768) return fail(c, SFM_ERR_CAPACITY, "extract: image side above 32768 pixels");
    if (o.n_octave_layers < 1 || o.n_octave_layers > 8) return fail(c, SFM_ERR_UNSUPPORTED, "extract: nOctaveLayers must be in 1..8");
    if (!(o.sigma > 0.0) || !(o.contrast_threshold >= 0.0) || !(o.edge_threshold > 0.0))
        return fail(c, SFM_ERR_INVALID, "extract: sigma / thresholds out of range");
    int max_kp = o.max_keypoints > 0 ? o.max_keypoints : SFM_MAX_ROWS - 1;
    if (max_kp >= SFM_MAX_ROWS) max_kp = SFM_MAX_ROWS - 1;
    CU_TRY(c, cudaSetDevice(c->device));
    if (!c->sift) c->sift = sift_workspace_create();
    SiftParams prm;
    prm.n_layers = o.n_octave_layers; prm.n_features = o.n_features > 0 ? o.n_features : 0; prm.contrast_threshold = o.contrast_threshold; prm.edge_threshold = o.edge_threshold;
    prm.sigma = o.sigma;
    int n = 0, launches = 0;
    std::string err;
    cudaError_t e = sift_extract(c->sift, gray, rows, cols, step_bytes, prm, max_kp, c->sm_count, c->stream, &n, &launches, c->feat_counts, &err);
    if (e != cudaSuccess)
        return fail(c, e == cudaErrorMemoryAllocation && !err.empty() && err.find("capacity") != std::string::npos ? SFM_ERR_CAPACITY : SFM_ERR_CUDA,
                    err.empty() ? cudaGetErrorString(e) : err);
    c->stat_launches += launches;
    c->stat_h2d += static_cast<int64_t>(rows) * cols;
    const int rc_app = append_features(c, sift_keypoints_device_raw(c->sift), sift_descriptors_device(c->sift), n, 128);
    if (rc_app != SFM_OK) return rc_app;
    if (n_keypoints) *n_keypoints = n;
    return SFM_OK;
}

void sfm_orb_opts_default(sfm_orb_opts* o) {
    if (!o) return;
    o->n_features = 500;             // cv::ORB::create defaults (features2d.hpp); the reference passes its feature-limit
    o->max_keypoints = 0;
    o->n_levels = 8; o->edge_threshold = 31; o->patch_size = 31; o->fast_threshold = 20;
    o->scale_factor = 1.2f;
    o->reserved = 0;
}

int sfm_features_extract_orb(sfm_ctx* c, const uint8_t* gray, int rows, int cols, size_t step_bytes, const sfm_orb_opts* opts,
                             int32_t* n_keypoints) {
    if (!c) return SFM_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    NvtxRange nvtx_range("sfm:features_extract_orb");
    sfm_orb_opts o;
    sfm_orb_opts_default(&o);
    if (opts) o = *opts;
    if (!gray || rows <= 0 || cols <= 0) return fail(c, SFM_ERR_INVALID, "extract: empty image");
    if (step_bytes == 0) step_bytes = static_cast<size_t>(cols);
    if (step_bytes < static_cast<size_t>(cols)) return fail(c, SFM_ERR_INVALID, "extract: step smaller than a row");
    if (rows > 32768 || cols > 32768) return fail(c, SFM_ERR_CAPACITY, "extract: image side above 32768 pixels");
    if (o.n_features < 0) return fail(c, SFM_ERR_INVALID, "extract: nfeatures must be >= 0");
    if (o.n_levels != 8 || o.edge_threshold != 31 || o.patch_size != 31 || o.fast_threshold != 20 || o.scale_factor != 1.2f)
        return fail(c, SFM_ERR_UNSUPPORTED, "extract: only cv::ORB::create's defaults (scaleFactor 1.2, 8 levels, edgeThreshold 31, patchSize 31, "
                                            "fastThreshold 20, HARRIS_SCORE, WTA_K 2) are built; the reference changes nfeatures only");
    // ties of retainBest can exceed nfeatures: capacity = 2 x nfeatures + slack, capped by the matcher's per-image limit
    int max_kp = o.max_keypoints > 0 ? o.max_keypoints : SFM_MAX_ROWS - 1;
    if (max_kp >= SFM_MAX_ROWS) max_kp = SFM_MAX_ROWS - 1;
    CU_TRY(c, cudaSetDevice(c->device));
    if (!c->orb) c->orb = orb_workspace_create();
    int n = 0, launches = 0;
    std::string err;
    cudaError_t e = orb_extract(c->orb, gray, rows, cols, step_bytes, o.n_features, max_kp, c->stream, &n, &launches, &err);
    if (e != cudaSuccess)
        return fail(c, e == cudaErrorMemoryAllocation && !err.empty() && err.find("capacity") != std::string::npos ? SFM_ERR_CAPACITY
                       : (e == cudaErrorInvalidValue && !err.empty() ? SFM_ERR_INVALID : SFM_ERR_CUDA),
                    err.empty() ? cudaGetErrorString(e) : err);
    c->stat_launches += launches;
    c->stat_h2d += static_cast<int64_t>(rows) * cols;
    const int rc_app = append_features(c, orb_keypoints_device_raw(c->orb), orb_descriptors_device(c->orb), n, 32);
    if (rc_app != SFM_OK) return rc_app;
    if (n_keypoints) *n_keypoints = n;
    return SFM_OK;
}

int sfm_features_orb_level(sfm_ctx* c, int what, int level, void* out, int32_t* width, int32_t* height) {
    if (!c) return SFM_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->orb) return fail(c, SFM_ERR_STATE, "features: no ORB extraction has run");
    const void* p = nullptr;
    int w = 0, h = 0;
    const int esz = orb_level_map(c->orb, what, level, &p, &w, &h);
    if (esz < 0) return fail(c, SFM_ERR_INVALID, "features: no such ORB level / map");
    if (width) *width = w;
    if (height) *height = h;
    if (!out) return SFM_OK;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaMemcpyAsync(out, p, static_cast<size_t>(w) * h * esz, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return SFM_OK;
}

int sfm_features_descriptor_bytes(const sfm_ctx* c, int* bytes) {
    if (!c || !bytes) return SFM_ERR_INVALID;
    *bytes = c->feat_cols;
    return SFM_OK;
}

int sfm_features_count(const sfm_ctx* c, int* n_images) {
    if (!c || !n_images) return SFM_ERR_INVALID;
    *n_images = static_cast<int>(c->feat_off.size()) - 1;
    return SFM_OK;
}

int sfm_features_last_counts(const sfm_ctx* c, int32_t counts[3]) {
    if (!c || !counts) return SFM_ERR_INVALID;
    for (int k = 0; k < 3; ++k) counts[k] = c->feat_counts[k];
    return SFM_OK;
}

int sfm_features_last_profile(const sfm_ctx* c, double* pyramid_ms, double* total_ms, double* pyramid_bytes) {
    if (!c) return SFM_ERR_INVALID;
    if (!c->sift) { if (pyramid_ms) *pyramid_ms = 0; if (total_ms) *total_ms = 0; if (pyramid_bytes) *pyramid_bytes = 0; return SFM_OK; }
    sift_last_profile(c->sift, pyramid_ms, total_ms, pyramid_bytes);
    return SFM_OK;
}

int sfm_features_download(sfm_ctx* c, int image, int32_t* n_keypoints, sfm_keypoint* kps, uint8_t* desc) {
    if (!c) return SFM_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (image < 0 || image + 1 >= static_cast<int>(c->feat_off.size())) return fail(c, SFM_ERR_INVALID, "features: image index out of range");
    const int64_t r0 = c->feat_off[image], n = c->feat_off[image + 1] - r0;
    if (n_keypoints) *n_keypoints = static_cast<int32_t>(n);
    if (n == 0 || (!kps && !desc)) return SFM_OK;
    CU_TRY(c, cudaSetDevice(c->device));
    if (kps) CU_TRY(c, cudaMemcpyAsync(kps, c->feat_kp.as<sfm_keypoint>() + r0, static_cast<size_t>(n) * sizeof(sfm_keypoint), cudaMemcpyDeviceToHost, c->stream));
    const size_t db = static_cast<size_t>(c->feat_cols);
    if (desc) CU_TRY(c, cudaMemcpyAsync(desc, c->feat_desc.as<uint8_t>() + r0 * db, static_cast<size_t>(n) * db, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->stat_d2h += n * ((kps ? 24 : 0) + (desc ? static_cast<int64_t>(db) : 0));
    return SFM_OK;
}

int sfm_features_pyramid_level(sfm_ctx* c, int octave, int level, float* out, int32_t* width, int32_t* height) {
    if (!c) return SFM_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->sift) return fail(c, SFM_ERR_STATE, "features: no extraction has run");
    int widths[kSiftMaxOctaves], heights[kSiftMaxOctaves], n_layers = 0;
    int64_t offs[kSiftMaxOctaves];
    const float* base = nullptr;
    const int n_oct = sift_pyramid_geometry(c->sift, &n_layers, widths, heights, offs, &base);
    if (octave < 0 || octave >= n_oct || level < 0 || level >= n_layers + 3) return fail(c, SFM_ERR_INVALID, "features: no such pyramid level");
    if (width) *width = widths[octave];
    if (height) *height = heights[octave];
    if (!out) return SFM_OK;
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t plane = static_cast<size_t>(widths[octave]) * heights[octave];
    CU_TRY(c, cudaMemcpyAsync(out, base + offs[octave] + level * plane, plane * 4, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return SFM_OK;
}

int sfm_bank_from_features(sfm_ctx* c) {
    if (!c) return SFM_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    NvtxRange nvtx_range("sfm:bank_from_features");
    const int n_images = static_cast<int>(c->feat_off.size()) - 1;
    CU_TRY(c, cudaSetDevice(c->device));
    c->run.valid = false;
    std::vector<int32_t> n_rows(std::max(n_images, 1));
    for (int i = 0; i < n_images; ++i) n_rows[i] = static_cast<int32_t>(c->feat_off[i + 1] - c->feat_off[i]);
    Bank& b = c->bank;
    const size_t db = static_cast<size_t>(c->feat_cols ? c->feat_cols : 128);        // 128: cv::SIFT rows, 32: cv::ORB rows
    int rc = bank_layout(c, b, n_images, n_rows.data(), static_cast<int>(db), SFM_CV_8U);
    if (rc != SFM_OK) return rc;
    CU_TRY(c, b.d_u8.ensure(std::max<size_t>(16, static_cast<size_t>(b.padded_rows) * db)));
    CU_TRY(c, b.d_kp.ensure(std::max<size_t>(16, static_cast<size_t>(b.padded_rows) * 8)));
    for (int i = 0; i < n_images; ++i) {
        if (n_rows[i] == 0) continue;
        CU_TRY(c, cudaMemcpyAsync(b.d_u8.as<uint8_t>() + static_cast<size_t>(b.row0[i]) * db,
                                  c->feat_desc.as<uint8_t>() + c->feat_off[i] * db, static_cast<size_t>(n_rows[i]) * db,
                                  cudaMemcpyDeviceToDevice, c->stream));
        CU_TRY(c, launch_keypoint_xy(c->feat_kp.as<sfm_keypoint>() + c->feat_off[i], n_rows[i], b.d_kp.as<float2>() + b.row0[i], c->stream));
        c->stat_launches++;
    }
    rc = bank_finish(c, b);
    if (rc != SFM_OK) return rc;
    b.have_kp = true;
    return SFM_OK;
}

int sfm_set_profiling(sfm_ctx* c, int on) {
    if (!c) return SFM_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    c->profiling = on != 0;
    return SFM_OK;
}

int sfm_last_profile(sfm_ctx* c, double* knn_ms, double* post_ms, int* knn_launches) {
    if (!c) return SFM_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    double k = 0, p = 0;
    if (c->prof_used > 0) CU_TRY(c, cudaEventSynchronize(c->prof_ev[c->prof_used - 1]));
    for (int i = 0; i + 3 <= c->prof_used; i += 3) {
        float a = 0, b2 = 0;
        CU_TRY(c, cudaEventElapsedTime(&a, c->prof_ev[i], c->prof_ev[i + 1]));
        CU_TRY(c, cudaEventElapsedTime(&b2, c->prof_ev[i + 1], c->prof_ev[i + 2]));
        k += a; p += b2;
    }
    if (knn_ms) *knn_ms = k;
    if (post_ms) *post_ms = p;
    if (knn_launches) *knn_launches = c->prof_used / 3 * (c->run.opts.cross_check ? 2 : 1);
    return SFM_OK;
}

int sfm_knn_match(sfm_ctx* c, const void* query, int nq, size_t q_step, const void* train, int nt, size_t t_step, int cols,
                  int cv_depth, int norm, int k, int engine, int32_t* nidx, float* dist) {
    if (!c) return SFM_ERR_INVALID;
    std::lock_guard<std::mutex> lk(c->mu);
    NvtxRange nvtx_range("sfm:knn_match");
    if (k != 1 && k != 2) return fail(c, SFM_ERR_UNSUPPORTED, "knnMatch: k must be 1 or 2");
    if (nq < 0 || nt < 0 || cols <= 0) return fail(c, SFM_ERR_INVALID, "knnMatch: bad shape");
    if (nq > 0 && (!nidx || !dist)) return fail(c, SFM_ERR_INVALID, "knnMatch: null output");
    CU_TRY(c, cudaSetDevice(c->device));
    Bank& b = c->scratch;
    const void* rows[2] = {query, train};
    const int32_t n_rows[2] = {nq, nt};
    const size_t esz = cv_depth == SFM_CV_32F ? 4 : 1;
    const size_t steps[2] = {q_step ? q_step : cols * esz, t_step ? t_step : cols * esz};
    int rc = bank_upload_host(c, b, 2, rows, n_rows, cols, steps, cv_depth);
    if (rc != SFM_OK) return rc;
    if (nq == 0) return SFM_OK;
    Engine eng;
    rc = pick_engine(c, b, norm, engine, &eng);
    if (rc != SFM_OK) return rc;
    cudaStream_t s = c->stream;
    const int rpu = rows_per_unit(eng);
    struct Meta { PairDesc pd; int64_t unit_prefix[2]; int64_t out_prefix[2]; } meta;
    meta.pd.q_row0 = 0; meta.pd.nq = nq; meta.pd.t_row0 = static_cast<int32_t>(b.row0[1]); meta.pd.nt = nt; meta.pd.out_row0 = 0;
    meta.unit_prefix[0] = 0; meta.unit_prefix[1] = (nq + rpu - 1) / rpu;
    meta.out_prefix[0] = 0; meta.out_prefix[1] = pad_rows(nq);
    CU_TRY(c, c->d_pairs.ensure(sizeof(Meta)));
    CU_TRY(c, cudaMemcpyAsync(c->d_pairs.p, &meta, sizeof(Meta), cudaMemcpyHostToDevice, s));
    CU_TRY(c, cudaStreamSynchronize(s));
    CU_TRY(c, c->slot[0].top2.ensure(static_cast<size_t>(pad_rows(nq)) * sizeof(Top2)));
    const PairDesc* d_pd = c->d_pairs.as<PairDesc>();
    const int64_t* d_unit = reinterpret_cast<const int64_t*>(c->d_pairs.as<uint8_t>() + offsetof(Meta, unit_prefix));
    if (eng == Engine::TF32) {
        CU_TRY(c, c->slot[0].aux.ensure(static_cast<size_t>(pad_rows(nq)) * 4));
        CU_TRY(c, cudaMemsetAsync(c->d_scalars.as<uint8_t>() + 32, 0, 16, s));
    }
    rc = launch_knn(c, b, eng, d_pd, d_unit, 1, meta.unit_prefix[1], c->slot[0].top2.as<Top2>(), c->slot[0].aux.as<float>());
    if (rc != SFM_OK) return rc;
    if (eng == Engine::TF32) {
        RefineF32Args fa;
        fa.blk_pair = nullptr;
        fa.top2 = c->slot[0].top2.as<Top2>(); fa.aux = c->slot[0].aux.as<float>(); fa.pairs = d_pd;
        fa.out_prefix = reinterpret_cast<const int64_t*>(c->d_pairs.as<uint8_t>() + offsetof(Meta, out_prefix));
        fa.n_pairs = 1; fa.staged_rows = pad_rows(nq); fa.bank = b.d_f32.as<float>(); fa.fnorm2 = b.d_fnorm.as<float>();
        fa.nb_max = b.f_nb_max; fa.all_rows = 1; fa.swap_roles = 0; fa.ratio = 0.0;
        fa.stats = reinterpret_cast<unsigned long long*>(c->d_scalars.as<uint8_t>() + 32);
        CU_TRY(c, launch_refine_f32(fa, s));
        c->stat_launches++;
    }
    if (eng == Engine::TC && k == 2) {
        RefineArgs ra{};
        ra.top2 = c->slot[0].top2.as<Top2>(); ra.pairs = d_pd;
        ra.out_prefix = reinterpret_cast<const int64_t*>(c->d_pairs.as<uint8_t>() + offsetof(Meta, out_prefix));
        ra.n_pairs = 1; ra.staged_rows = pad_rows(nq); ra.bank = b.d_u8.as<uint8_t>(); ra.norm2 = b.d_norm2.as<int32_t>();
        ra.all_rows = 1; ra.ratio = 0.0; ra.hamming = norm == SFM_NORM_HAMMING;
        CU_TRY(c, launch_refine_second(ra, s));
        c->stat_launches++;
    }
    const size_t out_bytes = static_cast<size_t>(nq) * k * 4;
    CU_TRY(c, c->d_knn.ensure(2 * out_bytes));
    int32_t* d_idx = c->d_knn.as<int32_t>();
    float* d_dist = reinterpret_cast<float*>(c->d_knn.as<uint8_t>() + out_bytes);
    CU_TRY(c, launch_top2_to_arrays(c->slot[0].top2.as<Top2>(), nq, k, norm, d_idx, d_dist, s));
    c->stat_launches++;
    CU_TRY(c, c->h_knn.ensure(2 * out_bytes));
    CU_TRY(c, cudaMemcpyAsync(c->h_knn.p, d_idx, 2 * out_bytes, cudaMemcpyDeviceToHost, s));
    CU_TRY(c, cudaStreamSynchronize(s));
    std::memcpy(nidx, c->h_knn.p, out_bytes);
    std::memcpy(dist, c->h_knn.as<uint8_t>() + out_bytes, out_bytes);
    c->stat_d2h += static_cast<int64_t>(2 * out_bytes);
    c->run.valid = false;
    return SFM_OK;
}

}  // extern "C"
