#include "pair_selector.h"

#include <stdexcept>

#include "../../include/sfmmatch.h"

namespace sfmhost {

PairList unorderedPairs(std::size_t n) {
    PairList out;
    if (n > 1) out.reserve(n * (n - 1) / 2);
    for (std::size_t l = 0; l < n; ++l)
        for (std::size_t r = l + 1; r < n; ++r) out.emplace_back(static_cast<int32_t>(l), static_cast<int32_t>(r));
    return out;
}

PairList videoPairs(std::size_t n, int sequenceLength) {
    if (sequenceLength < 2) throw std::invalid_argument("sequence length must not be smaller than 2 (self plus one successor)");
    PairList out;
    const std::size_t window = static_cast<std::size_t>(sequenceLength - 1);
    for (std::size_t l = 0; l < n; ++l)
        for (std::size_t r = l + 1; r < n && r - l <= window; ++r)
            out.emplace_back(static_cast<int32_t>(l), static_cast<int32_t>(r));
    return out;
}

PairList gridPairs(std::size_t n, int sequenceLength, int rowLength) {
    if (sequenceLength < 2) throw std::invalid_argument("sequence length must not be smaller than 2 (self plus one successor)");
    if (rowLength < 1) throw std::invalid_argument("row length must not be smaller than 1");
    // The reference computes ceil(size / rowLength) with an integer division, i.e. floor: shots of a trailing
    // partial row take part in no pair.  Shot i sits in cell (i / rowLength, i % rowLength).
    const std::size_t cols = static_cast<std::size_t>(rowLength), seq = static_cast<std::size_t>(sequenceLength);
    const std::size_t rows = n / cols;
    PairList out;
    for (std::size_t cell = 0; cell < rows * cols; ++cell) {
        const std::size_t r = cell / cols, c = cell % cols;
        // triangular "right/up" stencil, row offset outer, column offset inner
        for (std::size_t dr = 0; dr < seq && r + dr < rows; ++dr)
            for (std::size_t dc = (dr == 0 ? 1 : 0); dr + dc < seq && c + dc < cols; ++dc)
                out.emplace_back(static_cast<int32_t>(cell), static_cast<int32_t>((r + dr) * cols + (c + dc)));
    }
    return out;
}

PairList selectPairs(std::size_t n, int featureSequence, int featureGridLength) {
    if (featureSequence >= 2) {
        if (featureGridLength >= 1) return gridPairs(n, featureSequence, featureGridLength);
        return videoPairs(n, featureSequence);
    }
    return unorderedPairs(n);     // any other value: the reference warns and uses Unordered
}

}  // namespace sfmhost

extern "C" int sfm_select_pairs(int n_shots, int feature_sequence, int feature_gridlength, int32_t* pairs,
                                int64_t capacity, int64_t* n_pairs) {
    if (n_shots < 0 || !n_pairs) return SFM_ERR_INVALID;
    try {
        const sfmhost::PairList pl = sfmhost::selectPairs(static_cast<std::size_t>(n_shots), feature_sequence, feature_gridlength);
        *n_pairs = static_cast<int64_t>(pl.size());
        if (pairs)
            for (int64_t i = 0; i < capacity && i < *n_pairs; ++i) { pairs[2 * i] = pl[i].first; pairs[2 * i + 1] = pl[i].second; }
        return SFM_OK;
    } catch (const std::exception&) {
        return SFM_ERR_INVALID;
    }
}
