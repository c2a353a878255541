"""K-feat divides small integers as q0 = c*rn; r = fma(-q0, n, c); q = fma(r, rn, q0) with
rn = RN(1/n).  This must be BIT-IDENTICAL to IEEE c/n (the reference's double division) for
every operand pair the features can produce.  Checked exhaustively on the CPU with libm's
correctly rounded fma() (same semantics as the GPU's DFMA)."""
import os
import subprocess
import tempfile

SRC = r"""
#include <math.h>
#include <stdio.h>
int main(void) {
    long bad = 0, cases = 0;
    for (int n = 1; n <= 4096; n++) {
        volatile double rn = 1.0 / (double)n;
        for (int c = 0; c <= n + 8 && c <= 4096; c++) {
            double a = c, b = n, q0 = a * rn, r = fma(-q0, b, a), q1 = fma(r, rn, q0);
            cases++;
            if (q1 != a / b) bad++;
        }
    }
    printf("%ld %ld\n", cases, bad);
    return 0;
}
"""


def test_reciprocal_fma_division_is_exact_for_small_integers():
    d = tempfile.mkdtemp()
    c = os.path.join(d, "chk.c")
    open(c, "w").write(SRC)
    exe = os.path.join(d, "chk")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", exe, c, "-lm"], check=True)
    cases, bad = (int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split())
    assert cases > 8_000_000 and bad == 0
