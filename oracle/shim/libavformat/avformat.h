/* TEST INFRASTRUCTURE ONLY (oracle/). Minimal stand-in for <libavformat/avformat.h>
 * so that the reference's include/bliss.h:5 and its analyser sources compile
 * verbatim without FFmpeg. The reference relies on libav headers to pull in
 * the libc headers below (e.g. fabs at src/amplitude_sort.c:65). */
#ifndef ORACLE_SHIM_AVFORMAT_H
#define ORACLE_SHIM_AVFORMAT_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#define LIBAVUTIL_VERSION_MAJOR 58
void *av_malloc(size_t size);
void av_free(void *ptr);
#endif
