/* TEST INFRASTRUCTURE - not part of the product; only tests/, tools/fmm_vs_edt.py and the oracle import it.
 *
 * CPU restatement of scikit-fmm 2022.3.26 `skfmm.distance(phi, dx=1)` (the call at reference
 * scripts/utils/leaf_scorer.py:69; requirements.txt:5 pins the version) for a 2-D field, order 2, no mask, not periodic.
 * scikit-fmm is a third-party dependency that is absent from /root/reference and from this image, so this follows its
 * PUBLISHED algorithm (src/base_marcher.cpp, src/distance_marcher.cpp, src/heap.cpp of that release), restated from the
 * description below - it is NOT validated against the library itself ("parity unpinned" for the values it produces):
 *
 *   1. initalizeFrozen: cells with phi == 0 are frozen at distance 0.  A cell whose phi has the opposite sign of a
 *      4-neighbour's is frozen at the sub-cell distance to the interpolated zero crossing,
 *      d = sign(phi) / sqrt(sum over dims of 1 / min_j(dx * phi / (phi - phi_neighbour_j))^2).
 *   2. initalizeNarrow: every other cell with a frozen 4-neighbour gets a trial value (updatePoint) and enters a
 *      binary min-heap keyed by |value|.
 *   3. solve: pop the smallest trial cell, freeze it; recompute the trial value of its non-frozen 4-neighbours (push
 *      when Far, decrease/increase key when already Narrow); second-order stencil: if the neighbour in direction j is
 *      frozen, the cell two steps away in that direction is recomputed too when it is Narrow.
 *   4. updatePoint (order 2): per dimension take the frozen neighbour with the smaller |value| (value1); if the cell
 *      behind it is frozen too and passes the release's test ((value2 <= value1 and value1 >= 0) or (value2 >= value1
 *      and value1 <= 0) - note that a neighbour at exactly 0 lets ANY non-negative value behind it pass), use the
 *      second-order one-sided difference  tp = (4 value1 - value2) / 3  with coefficient 9/4, else the first-order
 *      one; solve  a t^2 + b t + (c - 1) = 0  and take the root away from the zero level (larger root for phi > 0).
 *      A negative discriminant (the two dimensions disagree by more than a cell) falls back to the first-order
 *      update with the smallest neighbour alone.
 *
 * The reference only consumes the arg-max of the field (leaf_scorer.py:71); tools/fmm_vs_edt.py measures how often it
 * differs from the arg-max of the exact Euclidean transform that oracle, golden generator and CUDA path use.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { FAR = 0, NARROW = 1, FROZEN = 2 };

long long fmm_negative_discriminants = 0;   /* events since the library was loaded (diagnostic) */

typedef struct {
    int H, W;
    const double* phi;
    double* dist;
    unsigned char* flag;
    /* binary min-heap on |value| with back pointers */
    int* heap;      /* heap slot -> cell */
    double* key;    /* heap slot -> key */
    int* pos;       /* cell -> heap slot, -1 when not in the heap */
    int n;
} Fmm;

static void heap_swap(Fmm* f, int a, int b) {
    const int ca = f->heap[a], cb = f->heap[b];
    const double ka = f->key[a], kb = f->key[b];
    f->heap[a] = cb; f->key[a] = kb; f->pos[cb] = a;
    f->heap[b] = ca; f->key[b] = ka; f->pos[ca] = b;
}
static void sift_up(Fmm* f, int i) {
    while (i > 0) {
        const int p = (i - 1) / 2;
        if (f->key[p] <= f->key[i]) break;
        heap_swap(f, p, i);
        i = p;
    }
}
static void sift_down(Fmm* f, int i) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < f->n && f->key[l] < f->key[m]) m = l;
        if (r < f->n && f->key[r] < f->key[m]) m = r;
        if (m == i) break;
        heap_swap(f, m, i);
        i = m;
    }
}
static void heap_push(Fmm* f, int cell, double k) {
    const int i = f->n++;
    f->heap[i] = cell; f->key[i] = k; f->pos[cell] = i;
    sift_up(f, i);
}
static void heap_set(Fmm* f, int cell, double k) {
    const int i = f->pos[cell];
    const double old = f->key[i];
    f->key[i] = k;
    if (k < old) sift_up(f, i); else sift_down(f, i);
}
static int heap_pop(Fmm* f) {
    const int cell = f->heap[0];
    f->n--;
    if (f->n > 0) {
        f->heap[0] = f->heap[f->n]; f->key[0] = f->key[f->n]; f->pos[f->heap[0]] = 0;
        sift_down(f, 0);
    }
    f->pos[cell] = -1;
    return cell;
}

/* neighbour of cell i, `step` cells along dimension dim (0 = rows, 1 = columns); -1 outside the grid */
static int nb(const Fmm* f, int i, int dim, int step) {
    const int y = i / f->W, x = i - y * f->W;
    if (dim == 0) {
        const int yy = y + step;
        return (yy < 0 || yy >= f->H) ? -1 : yy * f->W + x;
    }
    const int xx = x + step;
    return (xx < 0 || xx >= f->W) ? -1 : y * f->W + xx;
}

static double update_point(const Fmm* f, int i) {
    double a = 0, b = 0, c = 0;
    double first_order_min = HUGE_VAL;
    for (int dim = 0; dim < 2; ++dim) {
        double v1 = HUGE_VAL, v2 = HUGE_VAL;
        for (int j = -1; j < 2; j += 2) {
            const int n1 = nb(f, i, dim, j);
            if (n1 != -1 && f->flag[n1] == FROZEN && fabs(f->dist[n1]) < fabs(v1)) {
                v1 = f->dist[n1];
                const int n2 = nb(f, i, dim, 2 * j);
                if (n2 != -1 && f->flag[n2] == FROZEN &&
                    ((f->dist[n2] <= v1 && v1 >= 0) || (f->dist[n2] >= v1 && v1 <= 0)))
                    v2 = f->dist[n2];     /* (no reset otherwise: a second-order value of the other direction is kept) */
            }
        }
        if (v2 < HUGE_VAL) {
            const double tp = (1.0 / 3.0) * (4.0 * v1 - v2);
            a += 9.0 / 4.0; b -= 2.0 * 9.0 / 4.0 * tp; c += 9.0 / 4.0 * tp * tp;
        } else if (v1 < HUGE_VAL) {
            a += 1.0; b -= 2.0 * v1; c += v1 * v1;
        }
        if (v1 < HUGE_VAL && fabs(v1) < fabs(first_order_min)) first_order_min = v1;
    }
    c -= 1.0;
    const double det = b * b - 4.0 * a * c;
    if (det >= 0) {
        if (f->phi[i] > 2.2e-16) return (-b + sqrt(det)) / 2.0 / a;
        return (-b - sqrt(det)) / 2.0 / a;
    }
    /* The stencils of the two dimensions are inconsistent.  The release either raises here ("negative discriminant") or
     * falls back; the restatement counts the event and continues one cell further than the nearest frozen neighbour. */
    ++fmm_negative_discriminants;
    return f->phi[i] > 0 ? first_order_min + 1.0 : first_order_min - 1.0;
}

/* dist[H*W] <- signed distance to the zero level set of phi; returns 0, or -1 when phi has no zero level set */
int fmm_distance_2d(const double* phi, int H, int W, double* dist) {
    const int N = H * W;
    Fmm f;
    f.H = H; f.W = W; f.phi = phi; f.dist = dist; f.n = 0;
    f.flag = (unsigned char*)calloc((size_t)N, 1);
    f.heap = (int*)malloc(sizeof(int) * (size_t)N);
    f.key = (double*)malloc(sizeof(double) * (size_t)N);
    f.pos = (int*)malloc(sizeof(int) * (size_t)N);
    if (!f.flag || !f.heap || !f.key || !f.pos) return -2;
    for (int i = 0; i < N; ++i) { dist[i] = HUGE_VAL; f.pos[i] = -1; }
    int frozen = 0;
    for (int i = 0; i < N; ++i)
        if (phi[i] == 0.0) { f.flag[i] = FROZEN; dist[i] = 0.0; ++frozen; }
    for (int i = 0; i < N; ++i) {
        if (f.flag[i] != FAR) continue;
        double ld[2] = {0, 0};
        int borders = 0;
        for (int dim = 0; dim < 2; ++dim)
            for (int j = -1; j < 2; j += 2) {
                const int n1 = nb(&f, i, dim, j);
                if (n1 != -1 && phi[i] * phi[n1] < 0) {
                    borders = 1;
                    const double d = phi[i] / (phi[i] - phi[n1]);
                    if (ld[dim] == 0 || ld[dim] > d) ld[dim] = d;
                }
            }
        if (borders) {
            double dsum = 0;
            for (int dim = 0; dim < 2; ++dim)
                if (ld[dim] > 0) dsum += 1.0 / ld[dim] / ld[dim];
            dist[i] = phi[i] < 0 ? -sqrt(1.0 / dsum) : sqrt(1.0 / dsum);
            f.flag[i] = FROZEN; ++frozen;
        }
    }
    if (!frozen) { free(f.flag); free(f.heap); free(f.key); free(f.pos); return -1; }
    /* narrow band */
    for (int i = 0; i < N; ++i) {
        if (f.flag[i] != FAR) continue;
        for (int dim = 0; dim < 2 && f.flag[i] == FAR; ++dim)
            for (int j = -1; j < 2; j += 2) {
                const int n1 = nb(&f, i, dim, j);
                if (n1 != -1 && f.flag[n1] == FROZEN) {
                    const double d = update_point(&f, i);
                    dist[i] = d; f.flag[i] = NARROW;
                    heap_push(&f, i, fabs(d));
                    break;
                }
            }
    }
    /* march */
    while (f.n > 0) {
        const int addr = heap_pop(&f);
        f.flag[addr] = FROZEN;
        for (int dim = 0; dim < 2; ++dim)
            for (int j = -1; j < 2; j += 2) {
                const int n1 = nb(&f, addr, dim, j);
                if (n1 != -1 && f.flag[n1] != FROZEN) {
                    const double d = update_point(&f, n1);
                    if (d != 0.0) {
                        dist[n1] = d;
                        if (f.flag[n1] == NARROW) heap_set(&f, n1, fabs(d));
                        else { f.flag[n1] = NARROW; heap_push(&f, n1, fabs(d)); }
                    }
                }
                /* second-order stencil: the cell behind a frozen neighbour */
                if (n1 != -1 && f.flag[n1] == FROZEN) {
                    const int n2 = nb(&f, addr, dim, 2 * j);
                    if (n2 != -1 && f.flag[n2] == NARROW) {
                        const double d = update_point(&f, n2);
                        if (d != 0.0) { dist[n2] = d; heap_set(&f, n2, fabs(d)); }
                    }
                }
            }
    }
    free(f.flag); free(f.heap); free(f.key); free(f.pos);
    return 0;
}
