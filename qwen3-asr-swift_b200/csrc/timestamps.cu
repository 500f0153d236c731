// timestamps.cu — host-side integer fix-ups of the forced aligner.
//
//   TimestampCorrection.enforceMonotonicity / longestIncreasingSubsequencePositions
//       /root/reference/Sources/Qwen3ASR/TimestampCorrection.swift:15-145
//   Qwen3ForcedAligner.findTrailingPlateauStart   /root/reference/Sources/Qwen3ASR/ForcedAligner.swift:191-216
// Pure integer / comparison logic, restated step by step (the anchor walk included) so the results are identical to the reference's
// on any input; the reference's unit tests (Tests/Qwen3ASRTests/ForcedAlignerTests.swift:213-259, 441-492) are ported in
// tests/test_aligner.py.
#include <math.h>

#include <algorithm>
#include <vector>

#include "../../include/q3asr.h"

namespace {

// TimestampCorrection.swift:103-144: patience LIS (strictly increasing), positions in ascending order
std::vector<int> lis_positions(const int* arr, int n) {
    std::vector<int> out;
    if (n <= 0) return out;
    std::vector<int> tails, tail_idx, parent((size_t)n, -1);
    for (int i = 0; i < n; i++) {
        int lo = 0, hi = (int)tails.size();
        while (lo < hi) {
            const int mid = (lo + hi) / 2;
            if (tails[(size_t)mid] < arr[i]) lo = mid + 1; else hi = mid;
        }
        if (lo == (int)tails.size()) {
            tails.push_back(arr[i]);
            tail_idx.push_back(i);
        } else {
            tails[(size_t)lo] = arr[i];
            tail_idx[(size_t)lo] = i;
        }
        parent[(size_t)i] = lo > 0 ? tail_idx[(size_t)lo - 1] : -1;
    }
    for (int idx = tail_idx.back(); idx != -1; idx = parent[(size_t)idx]) out.push_back(idx);
    std::reverse(out.begin(), out.end());
    return out;
}

}  // namespace

extern "C" {

int q3asr_lis_positions(const int* values, int n, int* positions, int* count) {
    if (n < 0 || (n > 0 && values == nullptr) || count == nullptr) return Q3ASR_ERR_INVALID;
    try {
        const std::vector<int> p = lis_positions(values, n);
        *count = (int)p.size();
        if (positions)
            for (size_t i = 0; i < p.size(); i++) positions[i] = p[i];
        return Q3ASR_OK;
    } catch (const std::exception&) {  // host allocation failure: nothing may cross the C boundary
        return Q3ASR_ERR_NOMEM;
    }
}

static int enforce_monotonicity_impl(const int* raw, int n, int* corrected);

int q3asr_enforce_monotonicity(const int* raw, int n, int* corrected) {
    if (n < 0 || (n > 0 && (raw == nullptr || corrected == nullptr))) return Q3ASR_ERR_INVALID;
    try {
        return enforce_monotonicity_impl(raw, n, corrected);
    } catch (const std::exception&) {
        return Q3ASR_ERR_NOMEM;
    }
}

static int enforce_monotonicity_impl(const int* raw, int n, int* corrected) {
    for (int i = 0; i < n; i++) corrected[i] = raw[i];
    if (n <= 1) return Q3ASR_OK;                                             // :16
    const std::vector<int> anchors = lis_positions(raw, n);                  // :19-26 (anchor value = raw[pos])
    if ((int)anchors.size() == n) return Q3ASR_OK;                           // :29-31
    std::vector<char> is_anchor((size_t)n, 0);
    for (int p : anchors) is_anchor[(size_t)p] = 1;
    const int na = (int)anchors.size();
    int anchor_idx = 0;
    for (int i = 0; i < n; i++) {                                            // :38-88
        if (is_anchor[(size_t)i]) {
            for (int a = 0; a < na; a++)
                if (anchors[(size_t)a] == i) { anchor_idx = a; break; }
            continue;
        }
        int prev = -1, next = -1;                                            // indices into anchors
        if (anchor_idx < na && anchors[(size_t)anchor_idx] < i) prev = anchor_idx;
        else if (anchor_idx > 0) prev = anchor_idx - 1;
        int next_idx = anchor_idx;
        while (next_idx < na && anchors[(size_t)next_idx] <= i) next_idx++;
        if (next_idx < na) next = next_idx;
        if (prev >= 0 && next >= 0) {
            const int pp = anchors[(size_t)prev], np = anchors[(size_t)next];
            const int pv = raw[pp], nv = raw[np];
            if (np - pp <= 3) {
                corrected[i] = (i - pp) <= (np - i) ? pv : nv;
            } else {
                const float t = (float)(i - pp) / (float)(np - pp);
                corrected[i] = pv + (int)(t * (float)(nv - pv));             // Int(Float) truncates toward zero
            }
        } else if (prev >= 0) {
            corrected[i] = raw[anchors[(size_t)prev]];
        } else if (next >= 0) {
            corrected[i] = raw[anchors[(size_t)next]];
        }
    }
    for (int i = 1; i < n; i++)                                              // :91-95
        if (corrected[i] < corrected[i - 1]) corrected[i] = corrected[i - 1];
    return Q3ASR_OK;
}

int q3asr_trailing_plateau_start(const float* start_times, int n, float tolerance, int min_size) {
    if (n <= 0 || start_times == nullptr) return n < 0 ? 0 : n;
    if (n <= min_size) return n;                                             // ForcedAligner.swift:203
    int plateau = n;
    for (int i = n - 1; i >= 1; i--) {
        if (fabsf(start_times[i] - start_times[i - 1]) < tolerance) plateau = i - 1; else break;
    }
    return (n - plateau) >= min_size ? plateau : n;
}

}  // extern "C"
