/* ORACLE (test infrastructure only): small radix-2 FFTs standing in for libavutil/tx. */
#ifndef JT_ORC_FFT_H
#define JT_ORC_FFT_H
typedef struct { float re, im; } orc_cf;
typedef struct { double re, im; } orc_cd;
/* in-place forward (inverse=0, e^{-i..}) / unnormalised inverse (inverse=1) complex FFT, n power of 2 */
void orc_fft_f32(orc_cf *x, int n, int inverse);
void orc_fft_f64(orc_cd *x, int n, int inverse);
#endif
