// oracle/_ref glue for the bag-of-words transform: the reference's own DBoW2 (Thirdparty/DBoW2/DBoW2/*.cpp,
// TemplatedVocabulary.h, DUtils/Random.cpp, Timestamp.cpp) compiled UNMODIFIED against the opencv2/boost shims and
// driven the way Frame::ComputeBoW does (/root/reference/src/Frame.cc:739-747, Converter::toDescriptorVector
// src/Converter.cc:24-32).  TEST INFRASTRUCTURE, NOT PRODUCT.
#include <cstdint>
#include <string>
#include <vector>

#include "DBoW2/FORB.h"
#include "DBoW2/TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;   // include/ORBVocabulary.h:29

extern "C" {

void *ref_vocab_load_text(const char *path) {
    ORBVocabulary *v = new ORBVocabulary();
    if (!v->loadFromTextFile(path)) { delete v; return nullptr; }
    return v;
}
void ref_vocab_free(void *v) { delete (ORBVocabulary *)v; }
int ref_vocab_size(void *v) { return (int)((ORBVocabulary *)v)->size(); }

// desc: n rows of 32 bytes.  Outputs: BowVector as (ids, vals) in map order, FeatureVector as CSR in map order.
// Returns 0; *n_bow / *n_fv are the entry counts (arrays must hold n entries each, fv_off n+1).
int ref_bow_transform(void *vp, const uint8_t *desc, int n, int levelsup, uint32_t *bow_ids, double *bow_vals, int *n_bow,
                      uint32_t *fv_nodes, int32_t *fv_off, uint32_t *fv_idx, int *n_fv) {
    ORBVocabulary *voc = (ORBVocabulary *)vp;
    std::vector<cv::Mat> vDesc;
    vDesc.reserve(n);
    cv::Mat all(n > 0 ? n : 1, 32, CV_8U, (void *)desc, 32);
    for (int j = 0; j < n; ++j) vDesc.push_back(all.row(j));
    DBoW2::BowVector bv;
    DBoW2::FeatureVector fv;
    voc->transform(vDesc, bv, fv, levelsup);
    int i = 0;
    for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++i) { bow_ids[i] = it->first; bow_vals[i] = it->second; }
    *n_bow = i;
    int k = 0, o = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++k) {
        fv_nodes[k] = it->first;
        fv_off[k] = o;
        for (size_t j = 0; j < it->second.size(); ++j) fv_idx[o++] = it->second[j];
    }
    fv_off[k] = o;
    *n_fv = k;
    return 0;
}

// per-feature word id (TemplatedVocabulary::transform(const TDescriptor&), :1050-1062)
void ref_bow_words(void *vp, const uint8_t *desc, int n, uint32_t *word_ids) {
    ORBVocabulary *voc = (ORBVocabulary *)vp;
    cv::Mat all(n > 0 ? n : 1, 32, CV_8U, (void *)desc, 32);
    for (int j = 0; j < n; ++j) word_ids[j] = voc->transform(all.row(j));
}

double ref_bow_score(void *vp, const uint32_t *ids1, const double *v1, int n1, const uint32_t *ids2, const double *v2, int n2) {
    ORBVocabulary *voc = (ORBVocabulary *)vp;
    DBoW2::BowVector a, b;
    for (int i = 0; i < n1; ++i) a.insert(a.end(), DBoW2::BowVector::value_type(ids1[i], v1[i]));
    for (int i = 0; i < n2; ++i) b.insert(b.end(), DBoW2::BowVector::value_type(ids2[i], v2[i]));
    return voc->score(a, b);
}

}  // extern "C"
