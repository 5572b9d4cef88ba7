// vocabulary_adapter.h -- ORBVocabulary (DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>, O3/include/ORBVocabulary.h)
// as far as the hot path uses it: loadFromTextFile and transform(features, BowVector&, FeatureVector&, levelsup),
// the call of Frame::ComputeBoW / KeyFrame::ComputeBoW (O3/src/Frame.cc:784-789).  The descent of every descriptor
// down the tree runs on the GPU (dvm_vocabulary_transform); the two ordered maps are filled through the
// reference's own BowVector / FeatureVector methods (addWeight, addIfNotExist, normalize, addFeature), in feature
// order, so their contents -- including the order of the floating-point sums -- are the reference's.
#pragma once
#include <cmath>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "dvm_host.h"

namespace dvm_host {

class Vocabulary {
public:
    Vocabulary() = default;
    Vocabulary(const Vocabulary&) = delete;
    Vocabulary& operator=(const Vocabulary&) = delete;
    ~Vocabulary() { dvm_vocabulary_destroy(h_); }

    // DBoW2/TemplatedVocabulary.h:1211-1287: "k L scoring weighting", then "parent is_leaf d0 .. d31 weight" per node
    bool loadFromTextFile(const std::string& filename)
    {
        // C stdio + strtol/strtod: the file has 1.08 M lines of 35 numbers (145 MB for ORBvoc.txt)
        std::FILE* f = std::fopen(filename.c_str(), "r");
        if (!f) return false;
        struct Closer { std::FILE* f; ~Closer() { std::fclose(f); } } closer{ f };
        std::vector<char> line(1 << 12);
        if (!std::fgets(line.data(), static_cast<int>(line.size()), f)) return false;
        int n1 = -1, n2 = -1;
        if (std::sscanf(line.data(), "%d %d %d %d", &k_, &L_, &n1, &n2) != 4) return false;
        if (k_ < 0 || k_ > 20 || L_ < 1 || L_ > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) return false;
        scoring_ = n1; weighting_ = n2;
        std::vector<int32_t> parent(1, 0), word(1, -1);
        std::vector<uint8_t> desc(32, 0);
        std::vector<double> weight(1, 0.0);
        int nwords = 0;
        while (std::fgets(line.data(), static_cast<int>(line.size()), f)) {
            char* p = line.data();
            char* end = nullptr;
            const long pid = std::strtol(p, &end, 10);
            if (end == p) continue;                      // blank line
            p = end;
            const long leaf = std::strtol(p, &end, 10);
            p = end;
            const size_t nid = parent.size();
            if (pid < 0 || static_cast<size_t>(pid) >= nid) return false;
            parent.push_back(static_cast<int32_t>(pid));
            for (int i = 0; i < 32; i++) { desc.push_back(static_cast<uint8_t>(std::strtol(p, &end, 10))); p = end; }
            weight.push_back(std::strtod(p, &end));
            word.push_back(leaf > 0 ? nwords++ : -1);
        }
        // children in file order under every parent (m_nodes[pid].children.push_back(nid))
        const size_t n = parent.size();
        std::vector<int32_t> start(n + 1, 0), children(n > 0 ? n - 1 : 0);
        for (size_t i = 1; i < n; i++) start[parent[i] + 1]++;
        for (size_t i = 0; i < n; i++) start[i + 1] += start[i];
        std::vector<int32_t> cur(start.begin(), start.end() - 1);
        for (size_t i = 1; i < n; i++) children[cur[parent[i]]++] = static_cast<int32_t>(i);
        dvm_vocabulary_destroy(h_);
        h_ = nullptr;
        check(dvm_vocabulary_create(&h_, device_from_env(), static_cast<int>(n), start.data(), children.data(), desc.data(),
                                    weight.data(), word.data(), L_),
              "ORBVocabulary::loadFromTextFile");
        return true;
    }

    bool empty() const { return h_ == nullptr; }

    // DBoW2/TemplatedVocabulary.h:1025-1086; MatT rows are 1 x 32 CV_8U descriptors (Converter::toDescriptorVector)
    template <class MatT, class BowVectorT, class FeatureVectorT>
    void transform(const std::vector<MatT>& features, BowVectorT& v, FeatureVectorT& fv, int levelsup) const
    {
        v.clear();
        fv.clear();
        if (empty()) return;
        const int n = static_cast<int>(features.size());
        std::vector<uint8_t> desc(static_cast<size_t>(n) * 32);
        for (int i = 0; i < n; i++) std::memcpy(&desc[static_cast<size_t>(i) * 32], features[i].ptr(0), 32);
        std::vector<int32_t> word(n > 0 ? n : 1), nid(n > 0 ? n : 1);
        std::vector<double> w(n > 0 ? n : 1);
        check(dvm_vocabulary_transform(h_, desc.data(), n, levelsup, word.data(), w.data(), nid.data()), "ORBVocabulary::transform");
        const bool must = scoring_ != 5;            // DotProductScoring does not normalise (DBoW2/ScoringObject.h:76-91)
        const bool accumulate = weighting_ == 0 || weighting_ == 1;   // TF_IDF, TF
        for (int i = 0; i < n; i++) {
            if (!(w[i] > 0)) continue;               // stopped word
            if (accumulate) v.addWeight(word[i], w[i]);
            else v.addIfNotExist(word[i], w[i]);
            fv.addFeature(nid[i], i);
        }
        if (accumulate && !v.empty() && !must) {
            const double nd = static_cast<double>(v.size());
            for (auto& kv : v) kv.second /= nd;
        }
        if (must) v.normalize(scoring_ == 1 ? norm_l2<BowVectorT>() : norm_l1<BowVectorT>());
    }

    int k() const { return k_; }
    int L() const { return L_; }

private:
    // the LNorm enumerators of the caller's DBoW2 (L1 = 0, L2 = 1, DBoW2/BowVector.h:32) through normalize()'s parameter type
    template <class B> static auto norm_l1() { return first_arg_t<decltype(&B::normalize)>(0); }
    template <class B> static auto norm_l2() { return first_arg_t<decltype(&B::normalize)>(1); }
    template <class M> struct first_arg;
    template <class C, class A> struct first_arg<void (C::*)(A)> { typedef A type; };
    template <class M> using first_arg_t = typename first_arg<M>::type;

    dvm_vocabulary* h_ = nullptr;
    int k_ = 0, L_ = 0, scoring_ = 0, weighting_ = 0;
};

} // namespace dvm_host
