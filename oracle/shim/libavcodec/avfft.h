/* TEST INFRASTRUCTURE ONLY (oracle/). Declarations of the libavcodec real-FFT
 * API used at src/frequency_sort.c:22,28,65,83,134 of the reference. The
 * implementation (shim_fft.c) is our own FFT, NOT FFmpeg's. */
#ifndef ORACLE_SHIM_AVFFT_H
#define ORACLE_SHIM_AVFFT_H
typedef float FFTSample;
typedef struct RDFTContext RDFTContext;
enum RDFTransformType { DFT_R2C, IDFT_C2R, IDFT_R2C, DFT_C2R };
RDFTContext *av_rdft_init(int nbits, enum RDFTransformType trans);
void av_rdft_calc(RDFTContext *s, FFTSample *data);
void av_rdft_end(RDFTContext *s);
#endif
