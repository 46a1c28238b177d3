/* Exhaustive check of the quotient used by avg_distance_tile_kernel (pansim_b200/csrc/select.cuh):
 *     r = RN(1 / b);  q0 = RN(a * r);  e = a - q0 * b (one FMA, exact);  q = RN(q0 + e * r)
 * equals the IEEE quotient RN(a / b) for every pair of integers 0 <= a <= b <= bmax.
 * Build: gcc -O2 -mfma -o check_recip_division check_recip_division.c -lm
 * Run:   ./check_recip_division [bmax]      (default 131072 = AVG_RCP_MAX, ~16 s; tests/ run it to 4096) */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

int main(int argc, char **argv)
{
    const int bmax = argc > 1 ? atoi(argv[1]) : 131072;
    long long bad = 0, n = 0;
    for (int b = 1; b <= bmax; b++) {
        const double db = (double)b, r = 1.0 / db;
        for (int a = 0; a <= b; a++) {
            const double da = (double)a;
            const double q0 = da * r;
            const double e = fma(-q0, db, da);
            const double q = fma(e, r, q0);
            if (q != da / db) {
                if (bad < 10) printf("mismatch a=%d b=%d\n", a, b);
                bad++;
            }
            n++;
        }
    }
    printf("checked %lld pairs up to b = %d: %lld mismatches\n", n, bmax, bad);
    return bad != 0;
}
