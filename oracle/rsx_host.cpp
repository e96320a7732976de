// rsx_host.cpp — host instantiation of pywr_b200/csrc/rsx_core.cuh with a one-thread context.
// TEST INFRASTRUCTURE ONLY (see oracle/oracle_engine.py): it lets the revised-simplex core be checked against
// the dictionary oracle and HiGHS in a container without a GPU.  The product library instantiates the same
// core for the device only.
#include <stdlib.h>
#include <string.h>

#include "../pywr_b200/csrc/rsx_core.cuh"

namespace {
struct HostCtx {
    static constexpr int nl = 1;
    int tid = 0, nt = 1, lane = 0, warp = 0, nw = 1;
    void sync() {}
    bool leader() const { return true; }
    int atomic_inc(int *p) { return (*p)++; }
    double warp_sum(double v) { return v; }
    void argmax(double &, int &, int &) {}
    double min(double v) { return v; }
    int bor(int v) { return v; }
    void mark(int) {}
};
}  // namespace

extern "C" int rsx_host_run(int m, int n, const int *csr_ptr, const int *csr_col, const double *csr_val, const int *csc_ptr,
                            const int *csc_row, const double *csc_val, int cost_dynamic, double *W, int *head,
                            unsigned char *stat, double *d, int *info, const double *c, const double *lc, const double *uc,
                            const double *lo, const double *hi, double *x, int cold, int *pivots, int *refactors, int mode, int refresh_pivots) {
    // rows longer than 12 entries take the "long row" path of the core (PD_LONG of the device front end)
    int n_long = 0;
    for (int i = 0; i < m; i++) n_long += csr_ptr[i + 1] - csr_ptr[i] > 12;
    int *long_rows = (int *)malloc(sizeof(int) * (n_long + 1));
    n_long = 0;
    for (int i = 0; i < m; i++)
        if (csr_ptr[i + 1] - csr_ptr[i] > 12) long_rows[n_long++] = i;
    rsx::Lp P;
    P.m = m; P.n = n; P.ldw = (m + 1) & ~1;
    P.csr_ptr = csr_ptr; P.csr_col = csr_col; P.csr_val = csr_val;
    P.csc_ptr = csc_ptr; P.csc_row = csc_row; P.csc_val = csc_val;
    P.cost_dynamic = cost_dynamic;
    P.n_long = n_long; P.long_len = 12; P.long_rows = long_rows;
    P.refresh_pivots = refresh_pivots;
    rsx::Scen S;
    S.W = W; S.head = head; S.stat = stat; S.d = d; S.info = info;
    S.c = c; S.lc = lc; S.uc = uc; S.lo = lo; S.hi = hi; S.x = x; S.ld = 1; S.bt = nullptr; S.ldx = 1;
    unsigned char *buf = (unsigned char *)calloc(rsx::work_bytes(m, n, mode) + 64, 1);
    unsigned char *scratch = (unsigned char *)calloc(rsx::scratch_bytes(m, n, mode) + 64, 1);
    rsx::Work K;
    rsx::work_carve(K, buf, scratch, m, n, mode);
    HostCtx cx;
    int pv = 0, rf = 0;
    const int st = rsx::run_scenario(cx, P, S, K, cold, pv, rf);
    *pivots += pv; *refactors += rf;
    free(buf);
    free(scratch);
    free(long_rows);
    return st;
}
