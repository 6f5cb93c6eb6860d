// Host restatement of fmod_small (gnss-sdr-rs_b200/csrc/trk_kernels.cu): the tracking kernel replaces fmodf in the
// epoch-end phase updates (do_tracking.rs:240-242, 265-267) by one exact FMA from |x| once the integer quotient is
// right.  Checked bit for bit (value and sign) against glibc fmodf, including arguments next to exact multiples.
#include <math.h>
#include <fenv.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#pragma STDC FENV_ACCESS ON
static float add_rd(float a, float b){ volatile float r; fesetround(FE_DOWNWARD); r = a + b; fesetround(FE_TONEAREST); return r; }
static float fmod_small(float x, float y, float inv_y){
    const float ax = fabsf(x);
    volatile float q = ax * inv_y;
    volatile float t = add_rd(q, 8388608.0f);
    volatile float kf = t + (-8388608.0f);
    float r = fmaf(-kf, y, ax);
    if (r >= y) r = fmaf(-(kf + 1.0f), y, ax);
    else if (r < 0.f) r = fmaf(-(kf - 1.0f), y, ax);
    return copysignf(r, x);
}
int main(){
    const float twopi = 6.28318530717958647692f;
    uint64_t s = 88172645463325252ull; long bad = 0;
    for (long i = 0; i < 20000000L; i++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        float u = (float)((s >> 11) * (1.0 / 9007199254740992.0));
        float x, y, inv;
        int sel = i & 3;
        if (sel == 0) { x = (u - 0.3f) * 200.0f; y = twopi; inv = 0.15915494309189535f; }
        else if (sel == 1) { x = u * 4000.0f - 100.f; y = 1023.f; inv = 9.775171065493646e-4f; }
        else if (sel == 2) { int k = (int)(s & 63); x = nextafterf(twopi * k, (s & 64) ? 1e9f : -1e9f) ; if (s & 128) x = twopi * k; y = twopi; inv = 0.15915494309189535f; }
        else { int k = (int)(s & 255); x = 1023.f * k + ((int)((s >> 8) & 3) - 1) * 6.1035156e-05f * (k ? k : 1); y = 1023.f; inv = 9.775171065493646e-4f; }
        float a = fmodf(x, y), b = fmod_small(x, y, inv);
        if (!(a == b) || signbit(a) != signbit(b)) { if (bad < 10) printf("x=%.9g y=%.9g fmodf=%.9g mine=%.9g\n", x, y, a, b); bad++; }
    }
    printf("bad=%ld\n", bad);
    return 0;
}
