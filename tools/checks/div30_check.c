#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
int main(void) {
    const float lm = 30.0f;
    const float rcp = 1.0f / lm;
    uint32_t top; memcpy(&top, &lm, 4);
    unsigned long bad = 0, n = 0; float lastbad = 0;
    for (uint32_t b = 0; b <= top; ++b) {
        float x; memcpy(&x, &b, 4);
        float q0 = x * rcp;
        float r = fmaf(-q0, lm, x);
        float q = fmaf(r, rcp, q0);
        float ref = x / lm;
        if (q != ref) { ++bad; lastbad = x; }
        ++n;
    }
    printf("checked %lu values, %lu mismatches, largest mismatching x = %a (rcp=%a)\n", n, bad, lastbad, rcp);
    return 0;
}
