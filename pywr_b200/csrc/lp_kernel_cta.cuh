// lp_kernel_cta.cuh — one CTA per scenario, dictionary staged in shared memory.
//
// The variant BASELINE.json's north_star describes for networks that do not fit the register-resident
// kernels (lp_kernel.cuh): "one CTA per scenario, the basis factor and working vectors staged in shared
// memory, pricing and ratio tests done as warp-shuffle argmin/argmax reductions".  The dictionary T
// (m x n doubles, row stride padded to an odd count so that a warp reading one column of consecutive
// rows is bank-conflict free) plus all working vectors live in dynamic shared memory (up to 227 KB:
// m*n <= ~26,000, e.g. 200 x 128); thread t owns rows t, t+NT, ... and columns t, t+NT, ...
//
// Same algorithm, same pivoting rules and the same floating point order as oracle/lp_oracle.c
// (dict_solve / dict_pivot / dict_refactor / dict_reduced_costs) and as the register kernels.
// This file provides the per-timestep path (Solver.solve: cython_glpk.pyx:821-1095 / :1649-1918).
#pragma once
#include "lp_kernel.cuh"

namespace b200lp {

template <int NT>
struct CtaDict {
    static constexpr int NW = NT / 32;
    // shared memory
    double *T;  // [m][ldt]
    double *LO, *HI, *xn, *d, *dw, *pr, *x, *cost, *beta, *V;
    int *head, *nb, *nbstat, *vstat;
    double *red_key;  // [NW]
    int *red_idx, *red_aux;  // [NW]
    int m, n, ldt, tid, pivots;

    __host__ __device__ static size_t bytes(int m, int n, int n_val) {
        const int ldt = n | 1;
        const size_t doubles = (size_t)m * ldt + 2 * (size_t)(1 + n + m) + 6 * (size_t)n + (size_t)m + (size_t)n_val + K_COUNT + NW + 4;
        const size_t ints = (size_t)m + 2 * (size_t)n + (size_t)(n + m) + 2 * NW + 8;
        return doubles * sizeof(double) + ints * sizeof(int) + 32;
    }
    __device__ void carve(unsigned char *base, int m_, int n_, int n_val) {
        m = m_; n = n_; ldt = n_ | 1; tid = threadIdx.x; pivots = 0;
        double *dp = reinterpret_cast<double *>(base);
        T = dp; dp += (size_t)m * ldt;
        LO = dp; dp += 1 + n + m;
        HI = dp; dp += 1 + n + m;
        xn = dp; dp += n; d = dp; dp += n; dw = dp; dp += n; pr = dp; dp += n; x = dp; dp += n; cost = dp; dp += n;
        beta = dp; dp += m;
        V = dp; dp += n_val + K_COUNT;
        red_key = dp; dp += NW + 4;
        int *ip = reinterpret_cast<int *>(dp);
        head = ip; ip += m; nb = ip; ip += n; nbstat = ip; ip += n; vstat = ip; ip += n + m;
        red_idx = ip; ip += NW; red_aux = ip;
    }

    // ---- block reductions (uniform control flow; every thread receives the result) ----------
    __device__ void block_argmax(double &key, int &idx, int &aux) {
        group_argmax<32>(key, idx, aux, 0xffffffffu);
        __syncthreads();
        if ((tid & 31) == 0) { red_key[tid >> 5] = key; red_idx[tid >> 5] = idx; red_aux[tid >> 5] = aux; }
        __syncthreads();
        key = red_key[0]; idx = red_idx[0]; aux = red_aux[0];
        for (int w = 1; w < NW; w++) {
            const double k2 = red_key[w]; const int i2 = red_idx[w], a2 = red_aux[w];
            if ((i2 >= 0) && (idx < 0 || k2 > key || (k2 == key && i2 < idx))) { key = k2; idx = i2; aux = a2; }
        }
    }
    __device__ double block_min(double v) {
        v = group_min<32>(v, 0xffffffffu);
        __syncthreads();
        if ((tid & 31) == 0) red_key[tid >> 5] = v;
        __syncthreads();
        v = red_key[0];
        for (int w = 1; w < NW; w++) v = fmin(v, red_key[w]);
        return v;
    }
    __device__ unsigned block_or(unsigned bits) { return (unsigned)__syncthreads_or((int)bits); }

    __device__ void basic_values() {
        for (int i = tid; i < m; i += NT) {
            const double *row = T + (size_t)i * ldt;
            double acc = 0.0;
            for (int j = 0; j < n; j++) acc = fma(row[j], xn[j], acc);
            beta[i] = acc;
        }
        __syncthreads();
    }

    // exchange basic position r with non-basic position q (dict_pivot)
    __device__ void pivot(int r, int q, bool use_dw) {
        const double inv = 1.0 / T[(size_t)r * ldt + q];
        const double fd = d[q], fw = dw[q];
        for (int j = tid; j < n; j += NT) pr[j] = (j == q) ? inv : -T[(size_t)r * ldt + j] * inv;
        __syncthreads();
        for (int i = tid; i < m; i += NT) {
            double *row = T + (size_t)i * ldt;
            if (i == r) {
                for (int j = 0; j < n; j++) row[j] = pr[j];
            } else {
                const double f = row[q];
                for (int j = 0; j < n; j++) row[j] = (j == q) ? f * inv : fma(f, pr[j], row[j]);
            }
        }
        for (int j = tid; j < n; j += NT) {
            d[j] = (j == q) ? fd * inv : fma(fd, pr[j], d[j]);
            if (use_dw) dw[j] = (j == q) ? fw * inv : fma(fw, pr[j], dw[j]);
        }
        __syncthreads();
        if (tid == 0) { const int t = head[r]; head[r] = nb[q]; nb[q] = t; }
        pivots++;
        __syncthreads();
    }

    // d_j = c_nb[j] + sum_i c_head[i] T[i][j], i ascending: exactly the oracle's order
    __device__ void reduced_costs() {
        for (int j = tid; j < n; j += NT) {
            const int vj = nb[j];
            double acc = (vj >= 0 && vj < n) ? cost[vj] : 0.0;
            for (int i = 0; i < m; i++) {
                const int v = head[i];
                if (v < n) acc = fma(cost[v], T[(size_t)i * ldt + j], acc);
            }
            d[j] = acc;
        }
        __syncthreads();
    }

    __device__ void slack_basis() {
        for (int j = tid; j < n; j += NT) { nb[j] = j; nbstat[j] = VS_NL; }
        for (int i = tid; i < m; i += NT) head[i] = n + i;
        __syncthreads();
    }

    __device__ int refactor(const DevStructure &st) {
        for (int j = tid; j < n; j += NT) vstat[nb[j]] = nbstat[j];
        __syncthreads();
        for (int i = tid; i < m; i += NT) vstat[head[i]] = VS_BASIC;
        __syncthreads();
        for (int e = tid; e < m * n; e += NT) {
            const int i = e / n, j = e - i * n;
            T[(size_t)i * ldt + j] = __ldg(&st.A_dense[(size_t)j * st.rows_cap + i]);
        }
        __syncthreads();
        for (int k = tid; k < st.n_dyn_a; k += NT) {
            const int row = __ldg(&st.dyn_a_row[k]), col = __ldg(&st.dyn_a_col[k]);
            const double v = __ldg(&st.dyn_a_scale[k]) * V[__ldg(&st.dyn_a_idx[k])];
            T[(size_t)row * ldt + col] += v - __ldg(&st.dyn_a_static[k]);  // distinct (row, col) per entry
        }
        for (int i = tid; i < m; i += NT) head[i] = n + i;
        for (int j = tid; j < n; j += NT) { nb[j] = j; d[j] = 0.0; }
        __syncthreads();
        int ok = 0;
        for (int q = 0; q < n; q++) {
            if (vstat[q] != VS_BASIC) continue;
            double best = TOL_PIVOT; int bi = -1, aux = 0;
            for (int i = tid; i < m; i += NT) {
                const int v = head[i];
                if (v < n || vstat[v] == VS_BASIC) continue;
                const double a = fabs(T[(size_t)i * ldt + q]);
                if (a > best) { best = a; bi = i; }
            }
            if (bi < 0) best = -1.0;
            block_argmax(best, bi, aux);
            if (bi < 0) { ok = -1; break; }
            const int save = pivots;
            pivot(bi, q, false);
            pivots = save;
        }
        for (int j = tid; j < n; j += NT) nbstat[j] = vstat[nb[j]];
        __syncthreads();
        return ok;
    }

    __device__ int place_nonbasics() {
        int infeas = 0;
        for (int j = tid; j < n; j += NT) {
            const int v = nb[j];
            const double lo = LO[v + 1], hi = HI[v + 1], dj = d[j];
            int s = nbstat[j];
            if (lo == hi) s = VS_NS;
            else if (lo > -INFINITY && hi < INFINITY) {
                if (dj > TOL_DUAL) s = VS_NL;
                else if (dj < -TOL_DUAL) s = VS_NU;
                else if (s != VS_NL && s != VS_NU) s = VS_NL;
            } else if (lo > -INFINITY) { s = VS_NL; if (dj < -TOL_DUAL) infeas = 1; }
            else if (hi < INFINITY) { s = VS_NU; if (dj > TOL_DUAL) infeas = 1; }
            else { s = VS_NF; if (fabs(dj) > TOL_DUAL) infeas = 1; }
            nbstat[j] = s;
            xn[j] = (s == VS_NU) ? hi : (s == VS_NF ? 0.0 : lo);
        }
        return (int)block_or((unsigned)infeas);
    }

    __device__ bool eligible(int j, double a) const {
        const int s = nbstat[j];
        if (s == VS_NL) return a > TOL_PIVOT;
        if (s == VS_NU) return a < -TOL_PIVOT;
        if (s == VS_NF) return fabs(a) > TOL_PIVOT;
        return false;
    }

    // dict_solve
    __device__ int solve() {
        const int cap_bland = 20 * (m + n) + 50, cap_fail = 200 * (m + n) + 1000;
        int shifted = place_nonbasics();
        for (int j = tid; j < n; j += NT) {
            double w = d[j];
            if (shifted) {
                const int s = nbstat[j];
                if ((s == VS_NL && w < -TOL_DUAL) || (s == VS_NU && w > TOL_DUAL) || (s == VS_NF && fabs(w) > TOL_DUAL)) w = 0.0;
            }
            dw[j] = w;
        }
        __syncthreads();
        int phase = 0;
        for (int it = 0;; it++) {
            if (it > cap_fail) return B200LP_ST_ITERLIMIT;
            const bool bland = it > cap_bland;
            basic_values();
            int r, q, to_upper;
            if (phase == 0) {
                double worst = -1.0; int bi = -1, dir = 0;
                for (int i = tid; i < m; i += NT) {
                    const int v = head[i];
                    const double b = beta[i], lo = LO[v + 1], hi = HI[v + 1];
                    double viol = 0.0; int dd = 0;
                    if (b < lo - TOL_PRIMAL * (1.0 + fabs(lo))) { viol = lo - b; dd = +1; }
                    else if (b > hi + TOL_PRIMAL * (1.0 + fabs(hi))) { viol = b - hi; dd = -1; }
                    if (dd != 0) {
                        if (bland) viol = 1.0;
                        if (bi < 0 || viol > worst) { worst = viol; bi = i; dir = dd; }
                    }
                }
                block_argmax(worst, bi, dir);
                if (bi < 0) {
                    if (!shifted) return B200LP_ST_OPTIMAL;
                    phase = 1; shifted = 0;
                    continue;
                }
                r = bi;
                const double *prow = T + (size_t)r * ldt;
                double rmin = INFINITY;
                for (int j = tid; j < n; j += NT) {
                    const double a = prow[j] * (double)dir;
                    if (eligible(j, a)) rmin = fmin(rmin, fabs(dw[j]) / fabs(a));
                }
                rmin = block_min(rmin);
                if (rmin == INFINITY) return B200LP_ST_INFEASIBLE;
                const double thr = rmin * (1.0 + TIE_REL) + TIE_ABS;
                double amax = -1.0; int aux = 0;
                q = -1;
                for (int j = tid; j < n; j += NT) {
                    const double a = prow[j] * (double)dir;
                    if (!eligible(j, a)) continue;
                    if (fabs(dw[j]) / fabs(a) > thr) continue;
                    const double key = bland ? 1.0 : fabs(a);
                    if (q < 0 || key > amax) { amax = key; q = j; }
                }
                block_argmax(amax, q, aux);
                to_upper = dir < 0;
            } else {
                double worst = -1.0; int dir = 0;
                q = -1;
                for (int j = tid; j < n; j += NT) {
                    const int s = nbstat[j];
                    const double dj = d[j];
                    int dd = 0;
                    if ((s == VS_NL || s == VS_NF) && dj < -TOL_DUAL) dd = +1;
                    else if ((s == VS_NU || s == VS_NF) && dj > TOL_DUAL) dd = -1;
                    if (dd != 0) {
                        const double key = bland ? 1.0 : fabs(dj);
                        if (q < 0 || key > worst) { worst = key; q = j; dir = dd; }
                    }
                }
                block_argmax(worst, q, dir);
                if (q < 0) return B200LP_ST_OPTIMAL;
                const int vin = nb[q];
                const double inlo = LO[vin + 1], inhi = HI[vin + 1];
                double tmin = inhi - inlo;
                if (!(tmin < INFINITY)) tmin = INFINITY;
                double step_min = INFINITY;
                for (int i = tid; i < m; i += NT) {
                    const int v = head[i];
                    const double a = T[(size_t)i * ldt + q] * (double)dir, lo = LO[v + 1], hi = HI[v + 1];
                    double step = INFINITY;
                    if (a > TOL_PIVOT) { if (hi < INFINITY) step = (hi - beta[i]) / a; }
                    else if (a < -TOL_PIVOT) { if (lo > -INFINITY) step = (lo - beta[i]) / a; }
                    if (step < 0.0) step = 0.0;
                    step_min = fmin(step_min, step);
                }
                step_min = block_min(step_min);
                if (step_min == INFINITY && tmin == INFINITY) return B200LP_ST_UNBOUNDED;
                if (tmin <= step_min) {
                    __syncthreads();
                    if (tid == 0) { nbstat[q] = dir > 0 ? VS_NU : VS_NL; xn[q] = dir > 0 ? inhi : inlo; }
                    __syncthreads();
                    continue;
                }
                const double thr = step_min * (1.0 + TIE_REL) + TIE_ABS;
                double amax = -1.0; int bi = -1, up = 0;
                for (int i = tid; i < m; i += NT) {
                    const int v = head[i];
                    const double a = T[(size_t)i * ldt + q] * (double)dir, lo = LO[v + 1], hi = HI[v + 1];
                    double step = INFINITY; int u = 0;
                    if (a > TOL_PIVOT) { if (hi < INFINITY) { step = (hi - beta[i]) / a; u = 1; } }
                    else if (a < -TOL_PIVOT) { if (lo > -INFINITY) step = (lo - beta[i]) / a; }
                    if (step < 0.0) step = 0.0;
                    if (step > thr) continue;
                    const double key = bland ? 1.0 : fabs(a);
                    if (bi < 0 || key > amax) { amax = key; bi = i; up = u; }
                }
                block_argmax(amax, bi, up);
                r = bi;
                to_upper = up;
            }
            const int vout = head[r];
            const double vlo = LO[vout + 1], vhi = HI[vout + 1];
            __syncthreads();
            pivot(r, q, phase == 0);
            if (tid == 0) {
                nbstat[q] = (vlo == vhi) ? VS_NS : (to_upper ? VS_NU : VS_NL);
                xn[q] = to_upper ? vhi : vlo;
            }
            __syncthreads();
        }
    }
};

template <int NT>
__global__ void __launch_bounds__(NT, 1)
lp_step_kernel_cta(const DevStructure st, const DevState state, const int S, const double days,
                   const double *__restrict__ slots, double *__restrict__ node_flow, double *__restrict__ col_flow,
                   double *__restrict__ objective) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int sidx = blockIdx.x;
    const int m = st.m, n = st.n, tid = threadIdx.x;
    CtaDict<NT> D;
    D.carve(smem_raw, m, n, st.n_val);
    double *V = D.V;

    for (int k = tid; k < st.n_slots; k += NT) V[k] = slots[(size_t)k * S + sidx];
    for (int k = tid; k < st.n_consts; k += NT) V[st.n_slots + k] = __ldg(&st.consts[k]);
    for (int k = tid; k < st.n_storage; k += NT) V[st.vol_base + k] = state.vol[(size_t)k * S + sidx];
    for (int v = tid; v < 1 + n; v += NT) { D.LO[v] = v == 0 ? -INFINITY : 0.0; D.HI[v] = INFINITY; }
    __syncthreads();

    // (a5/a6) row bounds
    int flags = 0;
    for (int i = tid; i < m; i += NT) {
        const int cls = __ldg(&st.row_class[i]);
        double l, u;
        if (cls == ROWC_STORAGE) {
            const int k = __ldg(&st.row_a[i]);
            const double maxv = V[__ldg(&st.st_maxv_idx[k])], minv = V[__ldg(&st.st_minv_idx[k])];
            const double v = V[__ldg(&st.st_vol_idx[k])];
            const bool active = V[__ldg(&st.st_active_idx[k])] != 0.0;
            if (isnan(maxv) || isnan(minv) || isnan(v)) flags |= 1;
            if (maxv == minv) { l = 0.0; u = 0.0; }
            else if (!active) { l = -INFINITY; u = INFINITY; }
            else {
                l = -fmax(v - minv, 0.0) / days;
                u = fmax(maxv - v, 0.0) / days;
                type_bounds(l, u);
            }
        } else {
            l = V[__ldg(&st.row_a[i])]; u = V[__ldg(&st.row_b[i])];
            if (isnan(l) || isnan(u)) flags |= 1;
            type_bounds(l, u);
        }
        if (l > u) flags |= 2;
        D.LO[1 + n + i] = l; D.HI[1 + n + i] = u;
    }
    // (a3) objective
    for (int j = tid; j < n; j += NT) {
        double acc = 0.0;
        for (int k = __ldg(&st.cost_indptr[j]); k < __ldg(&st.cost_indptr[j + 1]); k++)
            acc += __ldg(&st.cost_w[k]) * V[__ldg(&st.cost_idx[k])];
        if (isnan(acc)) flags |= 1;
        if (st.cost_clip && fabs(acc) < 1e-8) acc = 0.0;
        D.cost[j] = acc;
    }
    for (int k = tid; k < st.n_dyn_a; k += NT)
        if (isnan(V[__ldg(&st.dyn_a_idx[k])])) flags |= 1;
    const unsigned all = D.block_or((unsigned)flags);

    int *meta = state.meta + (size_t)sidx * 4;
    int status = (all & 1u) ? B200LP_ST_NAN : ((all & 2u) ? B200LP_ST_INFEASIBLE : B200LP_ST_OPTIMAL);
    int refactors = 0;
    if (status == B200LP_ST_OPTIMAL) {
        int valid = meta[0], npiv = meta[1];
        double *Tg = state.T + (size_t)sidx * m * n;
        for (int i = tid; i < m; i += NT) D.head[i] = state.head[(size_t)sidx * m + i];
        for (int j = tid; j < n; j += NT) {
            D.nb[j] = state.nb[(size_t)sidx * n + j];
            D.nbstat[j] = state.nbstat[(size_t)sidx * n + j];
            D.d[j] = state.d[(size_t)sidx * n + j];
        }
        if (valid)
            for (int e = tid; e < m * n; e += NT) { const int i = e / n, j = e - i * n; D.T[(size_t)i * D.ldt + j] = Tg[e]; }
        __syncthreads();
        bool need_refactor = !valid || st.n_dyn_a > 0 || npiv >= REFACTOR_PIVOTS;
        const bool fresh0 = need_refactor;
        for (int attempt = 0; attempt < 3; attempt++) {
            if (attempt == 1 && fresh0) continue;
            if (attempt == 2) D.slack_basis();
            if (attempt > 0) need_refactor = true;
            if (need_refactor) {
                if (D.refactor(st) != 0) { D.slack_basis(); D.refactor(st); }
                D.reduced_costs();
                npiv = 0; refactors++; valid = 1;
            } else if (st.cost_dynamic) {
                D.reduced_costs();
            }
            const int p0 = D.pivots;
            status = D.solve();
            npiv += D.pivots - p0;
            if (status == B200LP_ST_OPTIMAL) break;
        }
        if (status != B200LP_ST_OPTIMAL) valid = 0;
        __syncthreads();
        if (D.pivots > 0 || refactors > 0) {
            for (int e = tid; e < m * n; e += NT) { const int i = e / n, j = e - i * n; Tg[e] = D.T[(size_t)i * D.ldt + j]; }
            for (int i = tid; i < m; i += NT) state.head[(size_t)sidx * m + i] = D.head[i];
        }
        for (int j = tid; j < n; j += NT) {
            state.nb[(size_t)sidx * n + j] = D.nb[j];
            state.nbstat[(size_t)sidx * n + j] = D.nbstat[j];
            state.d[(size_t)sidx * n + j] = D.d[j];
        }
        if (tid == 0) { meta[0] = valid; meta[1] = npiv; }
    }
    if (tid == 0) {
        meta[2] = status;
        if (D.pivots) atomicAdd(&state.counters[0], (unsigned long long)D.pivots);
        if (refactors) atomicAdd(&state.counters[1], (unsigned long long)refactors);
        if (status != B200LP_ST_OPTIMAL) {
            atomicAdd(&state.fail[0], 1);
            atomicMin(&state.fail[1], (sidx << 4) | status);
        }
    }
    if (status != B200LP_ST_OPTIMAL) return;

    // (a8) columns, objective, node flows (beta is current: solve() returned right after basic_values)
    for (int j = tid; j < n; j += NT) { const int v = D.nb[j]; if (v < n) D.x[v] = D.xn[j]; }
    for (int i = tid; i < m; i += NT) { const int v = D.head[i]; if (v < n) D.x[v] = D.beta[i]; }
    __syncthreads();
    if (col_flow)
        for (int j = tid; j < n; j += NT) col_flow[(size_t)j * S + sidx] = D.x[j];
    if (objective && tid == 0) {
        double acc = 0.0;
        for (int j = 0; j < n; j++) acc = fma(D.cost[j], D.x[j], acc);
        objective[sidx] = acc;
    }
    if (node_flow) {
        for (int v = tid; v < st.N; v += NT) {
            double acc = 0.0;
            for (int k = __ldg(&st.flow_indptr[v]); k < __ldg(&st.flow_indptr[v + 1]); k++)
                acc = fma(__ldg(&st.flow_w[k]), D.x[__ldg(&st.flow_col[k])], acc);
            node_flow[(size_t)v * S + sidx] = acc;
        }
    }
}

}  // namespace b200lp
