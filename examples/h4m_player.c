/*
 * h4m_player.c -- the reference program's main loop (h4m:2353-2547: open the file, walk the GOP
 * blocks, decode every video record, convert it to RGB) written against libhvqm4_b200.so, in the
 * reference's own language.  Two variants in one file:
 *
 *   default          the HVQM4Player* calls (container walk + buffer rotation inside the library)
 *   -DUSE_SDK_CALLS  the seven SDK entry points driven exactly like the reference's decode_video()
 *                    (h4m:2078-2138), with the caller's own malloc'ed frame buffers and rotation
 *
 * Instead of writing output/video_rgb_N.ppm it prints, per record, the frame type, the display
 * index (the N of that file name) and a 64-bit FNV-1a hash of the planar picture and of the RGB
 * payload, so that a test can compare the two variants with each other and with the oracle.
 *
 *   gcc -O2 -Iinclude examples/h4m_player.c -Lhvqm4_b200 -lhvqm4_b200 -Wl,-rpath,$PWD/hvqm4_b200 -o h4m_player
 *   ./h4m_player file.h4m
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hvqm4.h"

static uint64_t fnv1a(const uint8_t *p, size_t n)
{
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
    return h;
}

static uint8_t *read_file(const char *path, size_t *len)
{
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *buf = malloc((size_t)n + 8);
    if (buf && fread(buf, 1, (size_t)n, f) != (size_t)n) { free(buf); buf = NULL; }
    fclose(f);
    if (buf) memset(buf + n, 0, 8);
    *len = (size_t)n;
    return buf;
}

int main(int argc, char **argv)
{
    if (argc != 2) { fprintf(stderr, "usage: %s file.h4m\n", argv[0]); return 2; }
    size_t len;
    uint8_t *file = read_file(argv[1], &len);
    if (!file) { perror(argv[1]); return 1; }

#ifndef USE_SDK_CALLS
    HVQM4Player *pl = HVQM4PlayerOpen(file, len);
    if (!pl) { fprintf(stderr, "not an HVQM4 file, unsupported geometry, or no CUDA device\n"); return 1; }
    HVQM4FileInfo info;
    HVQM4PlayerInfo(pl, &info);
    const size_t frame_bytes = (size_t)info.width * info.height * 3 / 2, rgb_bytes = (size_t)info.width * info.height * 3;
    uint8_t *rgb = malloc(rgb_bytes);
    const uint8_t *yuv;
    uint32_t disp, type;
    int rc;
    while ((rc = HVQM4PlayerNextFrame(pl, &yuv, &disp, &type)) == 1)
    {
        if (HVQM4PlayerFrameRGB(pl, rgb) != HVQM4_OK) return 1;
        printf("%c %u %016llx %016llx\n", "?IPB"[type >> 4], disp, (unsigned long long)fnv1a(yuv, frame_bytes), (unsigned long long)fnv1a(rgb, rgb_bytes));
    }
    const uint32_t err = HVQM4PlayerErrors(pl);
    HVQM4PlayerClose(pl);
    free(rgb);
    free(file);
    return rc < 0 || err ? 1 : 0;
#else
    /* main(), h4m:2385-2419 */
    HVQM4FileInfo info;
    const int n = HVQM4ParseFile(file, len, &info, NULL, 0);
    if (n < 0) { fprintf(stderr, "not an HVQM4 file (%d)\n", n); return 1; }
    HVQM4FrameRef *frames = malloc(sizeof *frames * (size_t)(n ? n : 1));
    HVQM4ParseFile(file, len, &info, frames, n);
    VideoInfo vi = {(uint16_t)info.width, (uint16_t)info.height, (uint8_t)info.h_samp, (uint8_t)info.v_samp, file[0x3A]};
    SeqObj seq;
    HVQM4InitDecoder();
    HVQM4InitSeqObj(&seq, &vi);
    void *work = malloc(HVQM4BuffSize(&seq));
    HVQM4SetBuffer(&seq, work);
    HVQM4SetVersion(&seq, info.version);      /* the reference pokes state->padding[0] instead, h4m:2414-2417 */
    const size_t frame_bytes = (size_t)info.width * info.height * 3 / 2, rgb_bytes = (size_t)info.width * info.height * 3;
    uint8_t *past = calloc(1, frame_bytes), *present = calloc(1, frame_bytes), *future = calloc(1, frame_bytes), *rgb = malloc(rgb_bytes);
    uint32_t gop_start = 0, in_gop = 0;
    int cur_gop = 0;
    for (int i = 0; i < n; ++i)
    {   /* decode_video(), h4m:2078-2138 */
        const HVQM4FrameRef *f = &frames[i];
        if (f->gop != cur_gop) { gop_start += in_gop; in_gop = 0; cur_gop = f->gop; }
        ++in_gop;
        if (f->frame_type != 0x30) { uint8_t *t = past; past = future; future = t; }
        const uint8_t *pic = file + f->offset;
        switch (f->frame_type)
        {
        case 0x10: HVQM4DecodeIpic(&seq, pic, present); break;
        case 0x20: HVQM4DecodePpic(&seq, pic, present, past); break;
        default: HVQM4DecodeBpic(&seq, pic, present, past, future); break;
        }
        if (HVQM4ConvertRGB(&seq, present, rgb) != HVQM4_OK) return 1;     /* dumpRGB, h4m:2126 */
        printf("%c %u %016llx %016llx\n", "?IPB"[f->frame_type >> 4], gop_start + f->disp_id, (unsigned long long)fnv1a(present, frame_bytes),
               (unsigned long long)fnv1a(rgb, rgb_bytes));
        if (f->frame_type != 0x30) { uint8_t *t = present; present = future; future = t; }
    }
    const uint32_t err = HVQM4GetLastError(&seq);
    HVQM4ReleaseBuffer(&seq);
    free(work); free(past); free(present); free(future); free(rgb); free(frames); free(file);
    return err ? 1 : 0;
#endif
}
