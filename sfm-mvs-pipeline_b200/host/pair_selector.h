// pair_selector.h — image-pair lists of the three matching strategies of the reference.
//   Unordered  UnorderedFeatureMatchingStrategy.cpp:32-37
//   Video      VideoFeatureMatchingStrategy.cpp:43-48   (sequenceLength >= 2 enforced, :31-36)
//   Grid       GridFeatureMatchingStrategy.cpp:48-85    (sequenceLength >= 2, rowLength >= 1, :30-42)
// and the switch -> strategy mapping of PhotogrammetrieCli::configureFeatureMatcherStrategy
// (PhotogrammetrieCli.cpp:320-340).
#pragma once
#include <cstddef>
#include <cstdint>
#include <utility>
#include <vector>

namespace sfmhost {

using PairList = std::vector<std::pair<int32_t, int32_t>>;

// throw std::invalid_argument on bad parameters, like the reference's setters
PairList unorderedPairs(std::size_t nShots);
PairList videoPairs(std::size_t nShots, int sequenceLength);
PairList gridPairs(std::size_t nShots, int sequenceLength, int rowLength);
// feature-sequence / feature-gridlength switch values -> pair list
PairList selectPairs(std::size_t nShots, int featureSequence, int featureGridLength);

}  // namespace sfmhost
