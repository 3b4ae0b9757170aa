// k_umi_extract.cuh — BamUtil::getUMI(string qname, const string& prefix) (bamutil.cpp:40-112) for a batch of
// names, straight into the 4-bit UMI code the grouping kernel reads (gencore_b200.h "Encoding conventions").
// One thread per name: the scan is sequential in the reference (find_last_of, then a walk) and names are short.
#pragma once

#include "device_common.cuh"

namespace gcb {

constexpr int UMI_EXTRACT_THREADS = 128;
constexpr int UMI_MAX_PREFIX = 32;

struct UmiPrefix {  // the --umi_prefix string (options.h:15-61), by value in the launch parameters
    char s[UMI_MAX_PREFIX];
    int32_t len;
};

GCB_DEV bool is_umi_char(char c) { return c == 'A' || c == 'T' || c == 'C' || c == 'G' || c == '_'; }
GCB_DEV int umi_char_code(char c) { return c == 'A' ? 1 : c == 'C' ? 2 : c == 'G' ? 3 : c == 'T' ? 4 : c == '_' ? 5 : 0; }

// status per name
constexpr uint8_t UMI_OK = 0;        // code written (an empty UMI is all-zero words)
constexpr uint8_t UMI_TOO_LONG = 1;  // the UMI has more than 16*umi_words characters: code truncated, caller must widen

__global__ void __launch_bounds__(UMI_EXTRACT_THREADS) umi_extract_kernel(const char *names, const int64_t *name_off, int32_t n, UmiPrefix prefix,
                                                                          int32_t umi_words, uint64_t *out, uint8_t *status) {
    GCB_GRID_DEP();
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= n) return;
    const char *q = names + name_off[i];
    const int len = (int)(name_off[i + 1] - name_off[i]);
    int start = 0, ulen = 0;
    if (prefix.len > 0) {  // bamutil.cpp:45-63: find_last_of(prefix) is the last character that is ANY character of prefix
        int pos = -1;
        for (int k = len - 1; k >= 0 && pos < 0; k--) {
            const char c = q[k];
            for (int p = 0; p < prefix.len; p++)
                if (prefix.s[p] == c) pos = k;
        }
        if (pos >= 0) {
            start = pos + 2;
            for (int k = start; k < len && is_umi_char(q[k]); k++) ulen++;
        }
    } else {  // bamutil.cpp:65-111: the text after the last ':', one leading '_' skipped, only ATCG and at most one '_'
        int sep = len - 1;
        while (sep >= 0 && q[sep] != ':') sep--;
        if (sep >= 0 && sep < len - 1) {
            start = sep + 1;
            if (start < len - 1 && q[start] == '_') start++;
            int underscores = 0;
            bool good = true;
            for (int k = start; k < len && good; k++) {
                const char c = q[k];
                if (!is_umi_char(c)) good = false;
                if (c == '_' && ++underscores > 1) good = false;
            }
            if (good) ulen = len - start;
        }
    }
    uint8_t st = UMI_OK;
    if (ulen > 16 * umi_words) {
        ulen = 16 * umi_words;
        st = UMI_TOO_LONG;
    }
    uint64_t w = 0;
    uint64_t *o = out + (int64_t)i * umi_words;
    for (int k = 0; k < 16 * umi_words; k++) {
        if (k < ulen) w |= (uint64_t)umi_char_code(q[start + k]) << (60 - 4 * (k & 15));
        if ((k & 15) == 15) {
            o[k >> 4] = w;
            w = 0;
        }
    }
    status[i] = st;
}

}  // namespace gcb
