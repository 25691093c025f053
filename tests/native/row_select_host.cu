// Host harness for vsc2022_b200/csrc/row_select.cuh (CPU unit test of the per-thread
// top-K selection that the TN streaming kernel runs on the device).
#include <stdlib.h>

#include "row_select.cuh"

template <int K>
static int run(const float *rows, int n_rows, int lr, int *out_col, float *out_val) {
    float bm[vsc::kMaxRowBlocks], cv[vsc::kMaxCand];
    int cc[vsc::kMaxCand];
    int overflow = 0;
    for (int r = 0; r < n_rows; ++r) {
        float val[K]; int col[K];
        bool ok = vsc::select_row<K>(rows + (size_t)r * lr, lr, bm, cv, cc, 1, val, col);
        if (!ok) ++overflow;
        for (int i = 0; i < K; ++i) { out_col[r * K + i] = ok ? col[i] : -1; out_val[r * K + i] = val[i]; }
    }
    return overflow;
}

extern "C" int select_rows_host(const float *rows, int n_rows, int lr, int K, int *out_col, float *out_val) {
    switch (K) {
        case 1: return run<1>(rows, n_rows, lr, out_col, out_val);
        case 3: return run<3>(rows, n_rows, lr, out_col, out_val);
        case 5: return run<5>(rows, n_rows, lr, out_col, out_val);
        case 8: return run<8>(rows, n_rows, lr, out_col, out_val);
    }
    return -1;
}
