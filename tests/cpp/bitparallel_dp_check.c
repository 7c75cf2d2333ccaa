/* The derivation behind bv_align (dentist_b200/csrc/pile.cu), checked on the CPU: the bit-parallel unit-cost DP (Myers 1999 in
 * Hyyro's edit-distance form, 128-bit vertical-delta pairs per column) gives back EVERY cell of the plain DP through
 *   D[i][j] = j + popc(VP_j & low i bits) - popc(VN_j & low i bits),
 * and a traceback that rebuilds the three predecessor values from those columns takes the same direction as the cell DP at
 * every step (diagonal > A base unmatched > B base inserted among the minima).  That is why the consensus votes, the -B
 * bridges and the transposed trace points of the bit-parallel kernels equal the oracle's cell DP.  Test infrastructure. */
#include <stdio.h>
#include <stdlib.h>
typedef unsigned __int128 u128;
static int popc128(u128 x) { return __builtin_popcountll((unsigned long long)x) + __builtin_popcountll((unsigned long long)(x >> 64)); }
#define MASK(i) ((i) >= 128 ? ~(u128)0 : (((u128)1 << (i)) - 1))
static u128 VPs[252], VNs[252];
static int cell(int i, int j) { return j + popc128(VPs[j] & MASK(i)) - popc128(VNs[j] & MASK(i)); }

int main(int argc, char **argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 20000;
    srand(7);
    for (int iter = 0; iter < iters; iter++) {
        const int n = 1 + rand() % 128, m = rand() % 251;
        unsigned char a[128], b[256];
        for (int i = 0; i < n; i++) a[i] = rand() & 3;
        const int mode = rand() % 3;
        for (int j = 0; j < m; j++) b[j] = mode == 0 ? (rand() & 3) : ((j < n && rand() % 100 > 25) ? a[j] : (rand() & 3));
        static int D[129][252]; static unsigned char dir[129][252];
        for (int j = 0; j <= m; j++) D[0][j] = j;
        for (int i = 1; i <= n; i++) {
            D[i][0] = i;
            for (int j = 1; j <= m; j++) {
                const int d = D[i - 1][j - 1] + (a[i - 1] != b[j - 1]), u = D[i - 1][j] + 1, l = D[i][j - 1] + 1;
                int v = d; if (u < v) v = u; if (l < v) v = l;
                D[i][j] = v; dir[i][j] = (v == d) ? 0 : ((v == u) ? 1 : 2);
            }
        }
        u128 Peq[4] = {0, 0, 0, 0};
        for (int i = 0; i < n; i++) Peq[a[i]] |= (u128)1 << i;
        u128 VP = ~(u128)0, VN = 0; VPs[0] = VP; VNs[0] = VN;
        for (int j = 1; j <= m; j++) {
            const u128 Eq = Peq[b[j - 1]];
            const u128 D0 = (((Eq & VP) + VP) ^ VP) | Eq | VN;
            const u128 HP = VN | ~(D0 | VP), HN = D0 & VP;
            const u128 X = (HP << 1) | 1;                    /* row 0 grows by one per column: global alignment */
            VP = (HN << 1) | ~(D0 | X); VN = D0 & X;
            VPs[j] = VP; VNs[j] = VN;
        }
        for (int i = 0; i <= n; i++) for (int j = 0; j <= m; j++)
            if (cell(i, j) != D[i][j]) { printf("cell mismatch n=%d m=%d i=%d j=%d\n", n, m, i, j); return 1; }
        int i = n, j = m, c0 = cell(n, m);
        while (i > 0 || j > 0) {
            unsigned dr; int nc0;
            if (i == 0) { dr = 2; nc0 = j - 1; }
            else if (j == 0) { dr = 1; nc0 = i - 1; }
            else {
                const int up = c0 - ((int)((VPs[j] >> (i - 1)) & 1) - (int)((VNs[j] >> (i - 1)) & 1));
                const int left = cell(i, j - 1);
                const int dg = left - ((int)((VPs[j - 1] >> (i - 1)) & 1) - (int)((VNs[j - 1] >> (i - 1)) & 1));
                const int d = dg + (a[i - 1] != b[j - 1]), u = up + 1, l = left + 1;
                int v = d; if (u < v) v = u; if (l < v) v = l;
                if (v != c0) { printf("value mismatch\n"); return 1; }
                dr = (v == d) ? 0 : ((v == u) ? 1 : 2);
                if (dr != dir[i][j]) { printf("direction mismatch n=%d m=%d i=%d j=%d\n", n, m, i, j); return 1; }
                nc0 = dr == 0 ? dg : (dr == 1 ? up : left);
            }
            if (dr == 0) { i--; j--; } else if (dr == 1) i--; else j--;
            c0 = nc0;
        }
    }
    printf("all ok\n");
    return 0;
}
