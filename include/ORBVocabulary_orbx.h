/* ORBVocabulary_orbx.h — the two calls the tracking front-end makes on ORB_SLAM3::ORBVocabulary
 * (= DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>, include/ORBVocabulary.h:29), on the GPU over orbx.h:
 *   loadFromTextFile(path)                                    Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1338-1423
 *   transform(features, BowVector&, FeatureVector&, levelsup) :1127-1194, as Frame::ComputeBoW / KeyFrame::ComputeBoW
 *                                                             call it (src/Frame.cc:739-747)
 * SURVEY.md §8(f) rank 3.  The class is a template over the two result types so that this header does not need the
 * DBoW2 headers: instantiate it with DBoW2::BowVector (a std::map<WordId, WordValue>) and DBoW2::FeatureVector (a
 * std::map<NodeId, std::vector<unsigned int>>); any map types with those value types work.  Results are
 * bit-identical to DBoW2's, doubles included.  Scoring (loop closing) stays on the host: the vectors are tiny.
 */
#ifndef ORBVOCABULARY_ORBX_H
#define ORBVOCABULARY_ORBX_H

#include <string>
#include <vector>

#include "orbx.h"

namespace ORB_SLAM3
{

class ORBVocabularyDevice
{
public:
    explicit ORBVocabularyDevice(int device = 0): mDevice(device), mpVoc(nullptr) {}
    ~ORBVocabularyDevice() { orbx_vocab_destroy(mpVoc); }
    ORBVocabularyDevice(const ORBVocabularyDevice&) = delete;
    ORBVocabularyDevice& operator=(const ORBVocabularyDevice&) = delete;

    bool loadFromTextFile(const std::string &filename)
    {
        orbx_vocab_destroy(mpVoc);
        mpVoc = orbx_vocab_load_text(filename.c_str(), mDevice);
        return mpVoc != nullptr;
    }

    bool empty() const { return size() == 0; }
    unsigned int size() const
    {
        int nWords = 0;
        return (mpVoc && orbx_vocab_info(mpVoc, nullptr, nullptr, nullptr, &nWords) == ORBX_OK) ? (unsigned int)nWords : 0u;
    }

    // descriptors: n rows of 32 bytes (the cv::Mat the extractor filled; Converter::toDescriptorVector is not needed).
    template<class BowVector, class FeatureVector>
    bool transform(const unsigned char *descriptors, int n, BowVector &v, FeatureVector &fv, int levelsup) const
    {
        v.clear(); fv.clear();
        if(!mpVoc) return false;
        if(n <= 0) return true;
        std::vector<uint32_t> bowIds(n), fvNodes(n), fvIdx(n);
        std::vector<double> bowVals(n);
        std::vector<int32_t> fvOff(n + 1);
        int32_t nBow = 0, nFv = 0;
        if(orbx_bow_transform(mpVoc, descriptors, n, levelsup, nullptr, nullptr, bowIds.data(), bowVals.data(), &nBow,
                              fvNodes.data(), fvOff.data(), fvIdx.data(), &nFv) != ORBX_OK)
            return false;
        for(int i = 0; i < nBow; i++)            // already in map order: hinted insertion at the end is O(1)
            v.insert(v.end(), typename BowVector::value_type(bowIds[i], bowVals[i]));
        for(int k = 0; k < nFv; k++)
            fv.insert(fv.end(), typename FeatureVector::value_type(fvNodes[k],
                      typename FeatureVector::mapped_type(fvIdx.begin() + fvOff[k], fvIdx.begin() + fvOff[k + 1])));
        return true;
    }

    const char* LastError() const { return orbx_vocab_last_error(mpVoc); }

private:
    int mDevice;
    orbx_vocab *mpVoc;
};

} // namespace ORB_SLAM3

#endif
