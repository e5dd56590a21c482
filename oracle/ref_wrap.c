/*
 * oracle/ref_wrap.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin harness around the UNMODIFIED reference decoder.  The reference source is
 * not copied into this repository: it is #include'd by absolute path at build
 * time (see oracle/Makefile, -DHVQM4_REF_SRC=...), which is the only way to reach
 * its functions because every one of them is `static`
 * (/root/reference/h4m_audio_decode.c:275,819,828,957,1970,2018,2058) and its
 * main() is guarded by HVQM4_FFMPEG (h4m_audio_decode.c:2352).
 *
 * What this file adds on top of the reference:
 *   - an in-memory .h4m walker (the reference only has a FILE*-based main loop,
 *     h4m_audio_decode.c:2427-2537) that applies the same past/present/future
 *     rotation as decode_video() (h4m_audio_decode.c:2087-2093,2131-2137) and hands
 *     back the planar YUV `present` buffer instead of writing RGB PPMs;
 *   - per-section consumption counters, so the synthetic generator can be
 *     checked for "every section consumed to exactly its declared length";
 *   - direct entry points to a few static leaf operators for unit tests;
 *   - a fork()-based multi-process timing loop (the reference keeps decode state
 *     in globals, h4m_audio_decode.c:604-605, so threads are not an option).
 *
 * Output of the build goes to oracle/_ref/ (git-ignored, travels with gpurun).
 */
#define _GNU_SOURCE
#ifndef NATIVE
#define NATIVE 1
#endif
#define HVQM4_FFMPEG 1

#ifndef HVQM4_REF_SRC
#error "build with -DHVQM4_REF_SRC='\"/root/reference/h4m_audio_decode.c\"'"
#endif
#include HVQM4_REF_SRC

#include <time.h>
#include <unistd.h>
#include <sys/wait.h>

#define REF_API __attribute__((visibility("default")))

typedef struct RefStream
{
    const uint8_t *data;
    size_t len;
    size_t pos;
    HVQM4_header hdr;
    Player player;
    uint32_t picsize;
    uint32_t gops_left;
    uint32_t vid_left_in_gop;
    uint32_t aud_left_in_gop;
    uint32_t gop_start;       /* display index of first frame of current GOP */
    uint32_t vid_in_gop;
    int last_type;
    /* raw start pointers of the 17 sections of the last picture (for consumption checks) */
    const void *sec_start[17];
    uint32_t sec_size[17];
} RefStream;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

static int ref_inited;

REF_API RefStream *ref_open(const uint8_t *data, size_t len)
{
    if (len < 0x44)
        return NULL;
    if (!ref_inited)
    {
        HVQM4InitDecoder();
        ref_inited = 1;
    }
    RefStream *s = calloc(1, sizeof(*s));
    s->data = data;
    s->len = len;
    uint8_t raw[0x44];
    memcpy(raw, data, 0x44);
    load_header(&s->hdr, raw);
    s->pos = 0x44;
    s->gops_left = s->hdr.blocks;

    /* same call order as main(), h4m_audio_decode.c:2409-2419 */
    HVQM4InitSeqObj(&s->player.seqobj, (VideoInfo *)&s->hdr.hres);
    VideoState *state = calloc(1, HVQM4BuffSize(&s->player.seqobj));
    state->padding[0] = s->hdr.version == HVQM4_15;
    HVQM4SetBuffer(&s->player.seqobj, state);
    uint32_t res = s->player.seqobj.width * s->player.seqobj.height;
    uint32_t ss = s->player.seqobj.h_samp * s->player.seqobj.v_samp;
    s->picsize = (res * (ss + 2)) / ss;
    /* calloc instead of malloc: keeps the oracle deterministic even for streams
       that (illegally) reference a frame before it was decoded */
    s->player.past = calloc(1, s->picsize + 64);
    s->player.present = calloc(1, s->picsize + 64);
    s->player.future = calloc(1, s->picsize + 64);
    return s;
}

REF_API void ref_close(RefStream *s)
{
    if (!s)
        return;
    free(s->player.seqobj.state);
    free(s->player.past);
    free(s->player.present);
    free(s->player.future);
    free(s);
}

REF_API void ref_info(RefStream *s, int32_t out[6])
{
    out[0] = s->hdr.hres;
    out[1] = s->hdr.vres;
    out[2] = s->hdr.video_frames;
    out[3] = s->hdr.version == HVQM4_15 ? 15 : 13;
    out[4] = s->picsize;
    out[5] = s->hdr.blocks;
}

REF_API uint32_t ref_buffsize(int w, int h, int hs, int vs)
{
    SeqObj so;
    VideoInfo vi = {w, h, hs, vs, 0};
    HVQM4InitSeqObj(&so, &vi);
    return HVQM4BuffSize(&so);
}

static void record_sections(RefStream *s)
{
    VideoState *st = s->player.seqobj.state;
    BitBuffer *b[17] = {
        &st->basis_num[0].buf, &st->basis_num_run[0].buf, &st->basis_num[1].buf, &st->basis_num_run[1].buf,
        &st->dc_values[0].buf, &st->bufTree0[0].buf, &st->fixvl[0],
        &st->dc_values[1].buf, &st->bufTree0[1].buf, &st->fixvl[1],
        &st->dc_values[2].buf, &st->bufTree0[2].buf, &st->fixvl[2],
        NULL, NULL, NULL, NULL};
    if (s->last_type == I_FRAME)
    {
        b[13] = &st->dc_rle[0].buf; b[14] = &st->dc_rle[1].buf; b[15] = &st->dc_rle[2].buf;
    }
    else
    {
        b[13] = &st->mv_h.buf; b[14] = &st->mv_v.buf; b[15] = &st->mcb_type.buf; b[16] = &st->mcb_proc.buf;
    }
    for (int i = 0; i < 17; ++i)
    {
        s->sec_start[i] = b[i] ? b[i]->ptr : NULL;
        s->sec_size[i] = b[i] ? b[i]->size : 0;
    }
}

/*
 * Walks to the next video frame record, decodes it and copies the planar
 * Y|U|V present buffer to `out` (picsize bytes).  meta = {frame_type, disp_id,
 * display_index (gop_start + disp_id), record_size}.  Returns 1 on success,
 * 0 at end of stream, <0 on container error.
 */
REF_API int ref_decode_next(RefStream *s, uint8_t *out, int32_t meta[4])
{
    for (;;)
    {
        if (s->vid_left_in_gop == 0 && s->aud_left_in_gop == 0)
        {
            if (s->gops_left == 0)
                return 0;
            if (s->pos + 20 > s->len)
                return -1;
            /* GOP block header, h4m_audio_decode.c:2429-2438 */
            s->vid_left_in_gop = read32(s->data + s->pos + 8);
            s->aud_left_in_gop = read32(s->data + s->pos + 12);
            if (read32(s->data + s->pos + 16) != 0x01000000)
                return -2;
            s->pos += 20;
            s->gops_left--;
            s->gop_start += s->vid_in_gop;
            s->vid_in_gop = 0;
            continue;
        }
        if (s->pos + 8 > s->len)
            return -1;
        uint16_t id1 = read16(s->data + s->pos);
        uint16_t id2 = read16(s->data + s->pos + 2);
        uint32_t size = read32(s->data + s->pos + 4);
        s->pos += 8;
        if (s->pos + size > s->len)
            return -1;
        const uint8_t *rec = s->data + s->pos;
        s->pos += size;
        if (id1 == 0)
        {
            s->aud_left_in_gop--;
            continue;
        }
        if (id1 != 1)
            return -3;
        s->vid_left_in_gop--;
        s->vid_in_gop++;

        /* the reference mallocs size+3 because its word-wise bit reader overreads
           (h4m_audio_decode.c:2080-2082); copy so the slack exists at end of file too */
        uint8_t *frame = malloc(size + 8);
        memcpy(frame, rec, size);
        memset(frame + size, 0, 8);
        uint32_t disp_id = read32(frame);
        Player *pl = &s->player;
        if (id2 != B_FRAME)
        {
            void *t = pl->past; pl->past = pl->future; pl->future = t;
        }
        switch (id2)
        {
        case I_FRAME: HVQM4DecodeIpic(&pl->seqobj, frame + 4, pl->present); break;
        case P_FRAME: HVQM4DecodePpic(&pl->seqobj, frame + 4, pl->present, pl->past); break;
        case B_FRAME: HVQM4DecodeBpic(&pl->seqobj, frame + 4, pl->present, pl->past, pl->future); break;
        default: free(frame); return -4;
        }
        s->last_type = id2;
        record_sections(s);
        /* convert the section end pointers into consumed byte counts before `frame` dies */
        {
            const uint8_t *p = frame + 4 + 8;
            int nsec = id2 == I_FRAME ? 16 : 17;
            const uint8_t *dat = p + nsec * 4;
            for (int i = 0; i < 17; ++i)
            {
                if (i >= nsec || !s->sec_start[i]) { s->sec_start[i] = NULL; continue; }
                const uint8_t *sec = dat + read32(p + 4 * i) + 4;
                s->sec_start[i] = (const void *)(intptr_t)((const uint8_t *)s->sec_start[i] - sec);
            }
        }
        if (out)
            memcpy(out, pl->present, s->picsize);
        if (meta)
        {
            meta[0] = id2;
            meta[1] = disp_id;
            meta[2] = s->gop_start + disp_id;
            meta[3] = size;
        }
        free(frame);
        if (id2 != B_FRAME)
        {
            void *t = pl->present; pl->present = pl->future; pl->future = t;
        }
        return 1;
    }
}

/* consumed[i] = bytes the reference's reader advanced in section i of the last
   picture, size[i] = declared size (0 if the section is absent). */
REF_API void ref_section_usage(RefStream *s, int32_t consumed[17], int32_t size[17])
{
    for (int i = 0; i < 17; ++i)
    {
        consumed[i] = (int32_t)(intptr_t)s->sec_start[i];
        size[i] = s->sec_size[i];
    }
}

/* copy of the reference's work-buffer block maps after the last picture:
   for plane p, (h_blocks_safe*v_blocks_safe) {value,type} pairs incl. border */
REF_API int ref_get_map(RefStream *s, int plane, uint8_t *out, int32_t dims[2])
{
    HVQPlaneDesc *pl = &s->player.seqobj.state->planes[plane];
    dims[0] = pl->h_blocks_safe;
    dims[1] = pl->v_blocks_safe;
    if (out)
        memcpy(out, pl->border, (size_t)pl->h_blocks_safe * pl->v_blocks_safe * sizeof(BlockData));
    return 0;
}

REF_API void ref_get_nest(RefStream *s, uint8_t out[70 * 38])
{
    memcpy(out, s->player.seqobj.state->nest_data, 70 * 38);
}

/* Decode the whole stream `reps` times; returns seconds spent inside the
   HVQM4Decode?pic calls only (demux/rotation/memcpy are off the clock). */
typedef struct { const uint8_t *rec; uint32_t size; uint16_t type; } FrameRec;

static int index_frames(const uint8_t *data, size_t len, FrameRec **out)
{
    HVQM4_header hdr;
    uint8_t raw[0x44];
    memcpy(raw, data, 0x44);
    load_header(&hdr, raw);
    FrameRec *fr = malloc(sizeof(FrameRec) * (hdr.video_frames + 1));
    size_t pos = 0x44;
    int n = 0;
    for (uint32_t g = 0; g < hdr.blocks; ++g)
    {
        uint32_t nv = read32(data + pos + 8), na = read32(data + pos + 12);
        pos += 20;
        while (nv || na)
        {
            uint16_t id1 = read16(data + pos), id2 = read16(data + pos + 2);
            uint32_t size = read32(data + pos + 4);
            pos += 8;
            if (id1 == 1)
            {
                fr[n].rec = data + pos; fr[n].size = size; fr[n].type = id2; ++n; --nv;
            }
            else
                --na;
            pos += size;
        }
    }
    (void)len;
    *out = fr;
    return n;
}

static double bench_once(const uint8_t *data, size_t len, int reps, long *frames_out)
{
    RefStream *s = ref_open(data, len);
    FrameRec *fr;
    int n = index_frames(data, len, &fr);
    /* private padded copies of each record so the timed loop does no allocation */
    uint8_t **copies = malloc(sizeof(uint8_t *) * n);
    for (int i = 0; i < n; ++i)
    {
        copies[i] = calloc(1, fr[i].size + 8);
        memcpy(copies[i], fr[i].rec, fr[i].size);
    }
    Player *pl = &s->player;
    double t = 0;
    long frames = 0;
    for (int r = 0; r < reps; ++r)
    {
        for (int i = 0; i < n; ++i)
        {
            uint16_t ty = fr[i].type;
            if (ty != B_FRAME) { void *x = pl->past; pl->past = pl->future; pl->future = x; }
            double t0 = now_s();
            if (ty == I_FRAME) HVQM4DecodeIpic(&pl->seqobj, copies[i] + 4, pl->present);
            else if (ty == P_FRAME) HVQM4DecodePpic(&pl->seqobj, copies[i] + 4, pl->present, pl->past);
            else HVQM4DecodeBpic(&pl->seqobj, copies[i] + 4, pl->present, pl->past, pl->future);
            t += now_s() - t0;
            ++frames;
            if (ty != B_FRAME) { void *x = pl->present; pl->present = pl->future; pl->future = x; }
        }
    }
    for (int i = 0; i < n; ++i) free(copies[i]);
    free(copies);
    free(fr);
    ref_close(s);
    *frames_out = frames;
    return t;
}

REF_API double ref_bench(const uint8_t *data, size_t len, int reps, int64_t *frames)
{
    long f;
    double t = bench_once(data, len, reps, &f);
    *frames = f;
    return t;
}

/*
 * nproc forked workers each decode the stream `reps` times.  out[0] = wall seconds
 * (fork to last exit), out[1] = sum over workers of in-decode seconds, out[2] =
 * total frames decoded.  Returns 0 on success.
 */
REF_API int ref_bench_mp(const uint8_t *data, size_t len, int nproc, int reps, double out[3])
{
    int (*pipes)[2] = malloc(sizeof(int[2]) * nproc);
    pid_t *pids = malloc(sizeof(pid_t) * nproc);
    double t0 = now_s();
    for (int i = 0; i < nproc; ++i)
    {
        if (pipe(pipes[i])) return -1;
        pids[i] = fork();
        if (pids[i] == 0)
        {
            close(pipes[i][0]);
            long f;
            double t = bench_once(data, len, reps, &f);
            double msg[2] = {t, (double)f};
            if (write(pipes[i][1], msg, sizeof msg) != sizeof msg) _exit(1);
            _exit(0);
        }
        close(pipes[i][1]);
    }
    double tsum = 0, fsum = 0;
    int rc = 0;
    for (int i = 0; i < nproc; ++i)
    {
        double msg[2] = {0, 0};
        if (read(pipes[i][0], msg, sizeof msg) != sizeof msg) rc = -2;
        close(pipes[i][0]);
        int st;
        waitpid(pids[i], &st, 0);
        tsum += msg[0];
        fsum += msg[1];
    }
    out[0] = now_s() - t0;
    out[1] = tsum;
    out[2] = fsum;
    free(pipes);
    free(pids);
    return rc;
}

/*
 * The same over a SET of streams: worker i decodes streams i, i + nproc, i + 2 nproc, ... (cyclically) once each,
 * `reps` GOP decodes in all -- the bounded CPU sample of bench.py cycles the same distinct bitstreams as the GPU arm.
 */
REF_API int ref_bench_mp_streams(const uint8_t *const *datas, const size_t *lens, int n_streams, int nproc, int reps, double out[3])
{
    int (*pipes)[2] = malloc(sizeof(int[2]) * nproc);
    pid_t *pids = malloc(sizeof(pid_t) * nproc);
    double t0 = now_s();
    for (int i = 0; i < nproc; ++i)
    {
        if (pipe(pipes[i])) return -1;
        pids[i] = fork();
        if (pids[i] == 0)
        {
            close(pipes[i][0]);
            double msg[2] = {0, 0};
            for (int r = 0; r < reps; ++r)
            {
                const int k = (int)(((long)i + (long)r * nproc) % n_streams);
                long f;
                msg[0] += bench_once(datas[k], lens[k], 1, &f);
                msg[1] += (double)f;
            }
            if (write(pipes[i][1], msg, sizeof msg) != sizeof msg) _exit(1);
            _exit(0);
        }
        close(pipes[i][1]);
    }
    double tsum = 0, fsum = 0;
    int rc = 0;
    for (int i = 0; i < nproc; ++i)
    {
        double msg[2] = {0, 0};
        if (read(pipes[i][0], msg, sizeof msg) != sizeof msg) rc = -2;
        close(pipes[i][0]);
        int st;
        waitpid(pids[i], &st, 0);
        tsum += msg[0];
        fsum += msg[1];
    }
    out[0] = now_s() - t0;
    out[1] = tsum;
    out[2] = fsum;
    free(pipes);
    free(pids);
    return rc;
}

/* ---- leaf operators, exported for per-operator unit tests ---- */

REF_API void ref_WeightImBlock(uint8_t *dst, uint32_t stride, uint8_t v, uint8_t t, uint8_t b, uint8_t l, uint8_t r)
{
    WeightImBlock(dst, stride, v, t, b, l, r);
}

REF_API void ref_MotionComp4x4(uint8_t *dst, uint32_t dst_stride, const uint8_t *src, uint32_t src_stride, uint32_t hx, uint32_t hy)
{
    _MotionComp(dst, dst_stride, src, src_stride, hx, hy);
}

REF_API void ref_tables(int32_t div[16], int32_t mcdiv[512])
{
    if (!ref_inited) { HVQM4InitDecoder(); ref_inited = 1; }
    memcpy(div, divTable, sizeof divTable);
    memcpy(mcdiv, mcdivTable, sizeof mcdivTable);
}

/* ---- the reference's only observable output: dumpRGB (h4m:895-926), run on a caller-supplied frame.
   It writes a binary PPM; the pixel payload after the "P6\n<w> <h>\n255\n" header is returned. ---- */
REF_API int ref_yuv_to_rgb(const uint8_t *yuv, int w, int h, uint8_t *rgb)
{
    char path[64];
    snprintf(path, sizeof path, "/tmp/hvqm4_ref_rgb_%d.ppm", (int)getpid());
    Player pl;
    memset(&pl, 0, sizeof pl);
    pl.seqobj.width = (uint16_t)w;
    pl.seqobj.height = (uint16_t)h;
    pl.present = (void *)yuv;
    dumpRGB(&pl, path);
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    int pw = 0, ph = 0, maxv = 0, rc = -2;
    if (fscanf(f, "P6 %d %d %d", &pw, &ph, &maxv) == 3 && pw == w && ph == h && maxv == 255 && fgetc(f) == '\n')
        rc = fread(rgb, 1, (size_t)w * h * 3, f) == (size_t)w * h * 3 ? 0 : -3;
    fclose(f);
    unlink(path);
    return rc;
}

/* ---- the reference's audio decoder (decode_audio, h4m:185-258; its call is disabled upstream,
   h4m:2486-2507, but the function is compiled) run on caller-supplied bytes.  `state` = {hist, idx}
   per channel (int32 pairs) in and out; `data` starts BEHIND the sample count; returns the number
   of int16 values written to pcm (sample_count * channels), < 0 on failure. ---- */
REF_API int ref_decode_audio(int32_t *state, int channels, int first, uint32_t sample_count, const uint8_t *data, size_t len, int16_t *pcm)
{
    struct audio_state st;
    st.ch = calloc((size_t)channels, sizeof *st.ch);
    for (int c = 0; c < channels; ++c)
    {
        st.ch[c].hist = (int16_t)state[2 * c];
        st.ch[c].idx = (int8_t)state[2 * c + 1];
    }
    /* pad generously: the reference reads without bounds (the harness never asks for more samples than the bytes hold) */
    size_t padded = len + 16;
    uint8_t *copy = calloc(1, padded);
    memcpy(copy, data, len);
    FILE *in = fmemopen(copy, padded, "rb");
    char *out_buf = NULL;
    size_t out_len = 0;
    FILE *out = open_memstream(&out_buf, &out_len);
    int rc = -1;
    if (in && out)
    {
        decode_audio(&st, first, sample_count, in, out, channels);
        fflush(out);
        memcpy(pcm, out_buf, out_len);
        rc = (int)(out_len / 2);
    }
    if (in) fclose(in);
    if (out) fclose(out);
    free(out_buf);
    free(copy);
    for (int c = 0; c < channels; ++c)
    {
        state[2 * c] = st.ch[c].hist;
        state[2 * c + 1] = st.ch[c].idx;
    }
    free(st.ch);
    return rc;
}
