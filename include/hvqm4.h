/*
 * hvqm4.h -- C ABI of libhvqm4_b200.so, a B200-native HVQM4 1.3/1.5 picture decoder.
 *
 * PART 1 is the HVQM4 SDK picture-decode interface exactly as the reference exposes it
 * (the reference keeps these seven functions `static` in one translation unit and lists
 * them first in symbols.inc:2-8 as the SDK's public entry points); a program written
 * against the reference's API links against this library unchanged.  "h4m:N" below is
 * /root/reference/h4m_audio_decode.c line N.
 *
 * PART 2 holds the extensions the SDK lacks and a GPU decoder needs: version select,
 * error reporting, resource release, and a batched multi-stream interface that keeps
 * frame surfaces resident in device memory.
 *
 * Division of labour: the serial bitstream stage runs in C on host
 * threads and emits a per-picture symbol buffer into pinned memory; all pixel
 * reconstruction runs in CUDA kernels for sm_100a.  There is no CPU reconstruction
 * path: without a CUDA device every decode call fails (HVQM4_ERR_NO_DEVICE).
 */
#ifndef HVQM4_H
#define HVQM4_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ====================================================================== PART 1: SDK */

/* h4m:516-523.  `state` points into the caller's work buffer after HVQM4SetBuffer. */
typedef struct SeqObj
{
    void *state;
    uint16_t width;
    uint16_t height;
    uint8_t h_samp;   /* 2 = chroma halved horizontally */
    uint8_t v_samp;   /* 2 = chroma halved vertically */
} SeqObj;

/* h4m:533-540.  The reference passes a pointer into the parsed file header (h4m:2412). */
typedef struct VideoInfo
{
    uint16_t hres;
    uint16_t vres;
    uint8_t h_samp;
    uint8_t v_samp;
    uint8_t video_mode;
} VideoInfo;

/* replaces h4m:275 -- one-time global initialisation (constant tables; here also the
   CUDA context is created lazily on first decode). */
void HVQM4InitDecoder(void);

/* replaces h4m:819 -- copies geometry from the stream header into the sequence object. */
void HVQM4InitSeqObj(SeqObj *seqobj, VideoInfo *videoinfo);

/* replaces h4m:828 -- bytes of caller-owned work memory HVQM4SetBuffer needs.  (The
   value differs from the reference's: it is an ABI detail there as well, 32- vs 64-bit.) */
uint32_t HVQM4BuffSize(SeqObj *seqobj);

/* replaces h4m:957 -- binds the work buffer and builds per-stream decoder state.  The
   stream version defaults to 1.5; call HVQM4SetVersion for 1.3 streams (the reference
   pokes state->padding[0] before this call instead, h4m:2414-2417). */
void HVQM4SetBuffer(SeqObj *seqobj, void *workbuff);

/*
 * replace h4m:1970 / 2058 / 2018.  `frame` points 4 bytes into the video frame record,
 * just past the display id (h4m:2085,2100).  present/past/future are planar Y|U|V
 * buffers of width*height*3/2 bytes, stride = plane width (h4m:2343-2349); rotation of
 * the three buffers between calls is the caller's job (h4m:2087-2093, 2131-2137).
 *
 * Each pointer may be a host pointer (drop-in mode: references are uploaded when the
 * library has no device copy of them, the result is copied back before the call
 * returns) or a CUDA device pointer on the current device (zero-copy mode).
 */
void HVQM4DecodeIpic(SeqObj *seqobj, uint8_t const *frame, void *present);
void HVQM4DecodePpic(SeqObj *seqobj, uint8_t const *frame, void *present, void *past);
void HVQM4DecodeBpic(SeqObj *seqobj, uint8_t const *frame, void *present, void *past, void *future);

/* =============================================================== PART 2: extensions */

enum
{
    HVQM4_OK = 0,
    /* bits 0..7: stream errors found by the host stage (hvqm4_b200/csrc/symbuf.h SYM_ERR_*) */
    HVQM4_ERR_TRUNCATED  = 1 << 0,
    HVQM4_ERR_BAD_TREE   = 1 << 1,
    HVQM4_ERR_MCB_TYPE   = 1 << 2,
    HVQM4_ERR_MV_RANGE   = 1 << 3,
    HVQM4_ERR_PAIR_RANGE = 1 << 4,
    HVQM4_ERR_GEOMETRY   = 1 << 5,
    HVQM4_ERR_OVERFLOW   = 1 << 6,
    /* bits 16..: runtime errors */
    HVQM4_ERR_NO_DEVICE  = 1 << 16,  /* no CUDA device / driver: decoding is impossible */
    HVQM4_ERR_CUDA       = 1 << 17,  /* a CUDA call failed (see HVQM4GetLastCudaError) */
    HVQM4_ERR_ARGUMENT   = 1 << 18,
    HVQM4_ERR_NOMEM      = 1 << 19
};

/* version = 13 or 15.  Returns HVQM4_OK or HVQM4_ERR_ARGUMENT.  Call after HVQM4SetBuffer. */
int HVQM4SetVersion(SeqObj *seqobj, int version);

/* Optional: readable bytes behind the `frame` pointer of the NEXT decode call (record size
   minus 4).  The SDK protocol carries no length: without this call the sections are bounded by
   their own declared sizes and by what a picture of this geometry can need at most (64 frames'
   worth), so feed untrusted input with it. */
void HVQM4SetFrameBytes(SeqObj *seqobj, uint32_t bytes);

/* OR of all error bits since the last call; clears them.  The SDK entry points return void. */
uint32_t HVQM4GetLastError(SeqObj *seqobj);
int HVQM4GetLastCudaError(void);

/* Frees everything HVQM4SetBuffer allocated behind the work buffer (the SDK has no such call). */
void HVQM4ReleaseBuffer(SeqObj *seqobj);

/* Drop-in mode caches a device copy of every host frame buffer it has written, keyed by
   the host address.  A reference frame is fingerprinted (64 words spread over the buffer)
   before every use and uploaded again when the application has written into it; call this
   after a modification that such a sample may miss.  Device frame pointers (zero-copy mode)
   must be followed by 64 readable bytes: the half-sample filter reads whole aligned words. */
void HVQM4InvalidateFrame(SeqObj *seqobj, void *host_frame);

/* dumpRGB (h4m:895-926) of one frame through the GPU: `frame` is a planar picture of
   width * height * 3 / 2 bytes (host buffer: uploaded; device pointer: used in place); `rgb`
   (host) receives width * height * 3 bytes of interleaved R, G, B, identical to the reference's
   PPM payload.  Returns HVQM4_OK or HVQM4_ERR_* bits. */
int HVQM4ConvertRGB(SeqObj *seqobj, const void *frame, void *rgb);

/* ---------------------------------------------------------------- batched decoding
 * A batch is a pool of `n_streams` independent streams of identical geometry living on
 * one GPU: four device-resident frame surfaces per stream (past/present/future, rotated by
 * the library with the reference's rule, plus a spare that consecutive B pictures alternate
 * with so that a picture can be read back while the next one is reconstructed), per-stream
 * entropy state, a pool of host threads, and a ring of pinned/device staging arenas so that
 * the entropy decode and upload of the following steps overlap the reconstruction and
 * read-back of step k.
 */
typedef struct HVQM4Batch HVQM4Batch;

/* device < 0: current device.  host_threads <= 0: one per online CPU (capped at 64). */
HVQM4Batch *HVQM4BatchCreate(int device, int n_streams, int width, int height, int version, int host_threads);
void HVQM4BatchDestroy(HVQM4Batch *b);

/*
 * One step: decodes one picture for each listed stream (each stream at most once per
 * step).  frame_types[i] is 0x10 (I), 0x20 (P) or 0x30 (B) (h4m:2065-2070); frames[i]
 * points 4 bytes into the record like the SDK calls; frame_bytes[i] is the number of
 * readable bytes there.  Returns after the kernels are enqueued (asynchronous); returns
 * an OR of error bits.
 */
int HVQM4BatchDecode(HVQM4Batch *b, int n, const int32_t *stream_ids, const int32_t *frame_types,
                     const uint8_t *const *frames, const uint32_t *frame_bytes);

/*
 * Where the serial bitstream stage of this batch runs: 0 = host threads (default), 1 = on the GPU
 * (one warp per picture runs the same parser, compiled as device code; only raw picture bytes are
 * uploaded, symbol buffers never exist on the host).  Both give identical pictures.  The switch
 * must be made before the first HVQM4BatchDecode of the batch: the two stages keep separate
 * per-stream state, and a switch after the first picture is refused (HVQM4_ERR_ARGUMENT).  If the
 * device buffers of the GPU stage cannot be allocated the call returns HVQM4_ERR_NOMEM / _CUDA,
 * the batch stays on the host stage and the call may be repeated.  Recording (HVQM4BatchRecord)
 * needs the host stage.
 */
int HVQM4BatchSetEntropyMode(HVQM4Batch *b, int gpu);

/*
 * Extension of the mode above (it has no effect on the host stage): the pictures of streams 0 .. n_streams - 1 are parsed
 * by the batch's host threads -- the host stage of the same parser, writing their symbol buffers into the step's pinned
 * staging arena -- while the parse kernel takes the other streams of the step.  Both stages then feed ONE reconstruction
 * launch.  The GPU stage is the slower of the two stages of a dense end-to-end step and the host
 * cores are otherwise idle in this mode: a share of about one picture in eight per 16 host threads moves the step to the
 * PCIe bound.  Like the mode itself the share is fixed with the first picture (a stream's maps and nest live with the
 * stage that parsed its earlier pictures): a different value afterwards returns HVQM4_ERR_ARGUMENT.  Default 0.
 */
int HVQM4BatchSetHostShare(HVQM4Batch *b, int n_streams);

/* Diagnostics: SM cycles the GPU bitstream stage spent per phase since the last call, summed over
   pictures: header+trees, pass 1 (maps), record planning, pass 2 (scheduling), map copies, flat
   section decode, record fill, unused. */
void HVQM4DevEntropyProfile(uint64_t out[8]);

/* Waits for all enqueued work; returns accumulated error bits and clears them. */
int HVQM4BatchSync(HVQM4Batch *b);

/* Copies the most recently decoded picture of a stream (width*height*3/2 bytes) to host
   memory.  Synchronous. */
int HVQM4BatchReadFrame(HVQM4Batch *b, int stream_id, void *host_dst);

/* Asynchronous variant for pipelines: enqueues the device->host copy (host_dst should be
   pinned, see HVQM4HostAlloc) behind the step's kernel; complete after HVQM4BatchSync. */
int HVQM4BatchReadFrameAsync(HVQM4Batch *b, int stream_id, void *host_dst);

/* Same for n streams: frame of stream_ids[i] goes to host_base + i * host_stride. */
int HVQM4BatchReadFramesAsync(HVQM4Batch *b, int n, const int32_t *stream_ids, void *host_base, size_t host_stride);

/*
 * The reference's only observable output is every decoded frame converted to RGB (dumpRGB,
 * h4m:895-926, called from decode_video at h4m:2126: JPEG matrix in float, chroma replicated,
 * truncation, clamp).  Same result, bit for bit, from a conversion kernel on the GPU: the last
 * decoded picture of stream_ids[i] arrives at host_base + i * host_stride as width * height * 3
 * bytes of interleaved R, G, B.
 */
int HVQM4BatchReadFramesRGBAsync(HVQM4Batch *b, int n, const int32_t *stream_ids, void *host_base, size_t host_stride);

/* Device pointer of the most recently decoded picture of a stream (zero-copy consumers). */
void *HVQM4BatchFramePtr(HVQM4Batch *b, int stream_id);

/*
 * Reconstruction-only replay, used for kernel measurements: HVQM4BatchDecode with
 * `keep` steps recorded leaves each step's symbol buffers resident in device memory;
 * HVQM4BatchReplay re-launches the reconstruction kernels of the recorded steps in
 * order (no host stage, no upload) and returns the time the kernels took on the GPU,
 * measured with CUDA events on the launching stream, in milliseconds (< 0 on error).
 */
int HVQM4BatchRecord(HVQM4Batch *b, int enable);
float HVQM4BatchReplay(HVQM4Batch *b, int repeats);

/* Counters since creation: out[0] pictures, out[1] kernel launches, out[2] symbol bytes
   uploaded, out[3] algorithmic bytes (frame bytes written + reference bytes predicted from
   + symbol bytes), out[4] host-stage nanoseconds summed over threads, out[5] inter-coded
   macroblocks, out[6] total macroblocks, out[7] launches of the fused band kernel (process-wide). */
void HVQM4BatchStats(HVQM4Batch *b, uint64_t out[8]);

/* Number of reconstruction kernel launches issued by this process so far (all batches and
   SDK-mode decodes); used by benchmarks to report how much work really ran on the GPU. */
long long HVQM4KernelLaunches(void);

/* Reconstruction schedule: 0 = chosen by batch size (default); 1..4 = always the fused per-band
   kernel (one launch per step; 2, 3 or 4 also pins its CTAs per SM, 1 leaves that to the grid
   size); 5 = always the sweep kernel (one persistent CTA per SM walks a picture out of shared
   memory, reference rows and symbol slices staged by TMA bulk copies; pictures its plan does not
   serve and picture sizes it does not serve fall to the band kernel); 6 = always the row kernel
   (batch mode only: one persistent CTA per SM walks macroblock rows through a shared-memory
   pipeline, one reference patch per inter macroblock fetched by TMA tensor copies, output rows
   assembled in shared memory and written by bulk stores; pictures with predictions that leave
   their plane fall to the band kernel); 7 = the band kernel with the band assembled in shared
   memory and written by bulk stores of whole picture rows (two CTAs per SM; pictures too wide for
   that use the plain band kernel); < 0 = always the map kernel + record kernel pair.  All give identical pictures; the switch exists for tests and
   measurements.  Process-wide. */
void HVQM4SetReconMode(int mode);
/* Diagnostics of the sweep kernel: launches so far; nonzero if one of its CTAs ever gave up waiting on its
   copy pipeline (never on a healthy device; synchronises the device). */
long long HVQM4SweepLaunches(void);
int HVQM4SweepErrors(void);
/* the same for the row kernel */
long long HVQM4RowLaunches(void);
int HVQM4RowErrors(void);

/* Pinned host memory helpers (cudaHostAlloc / cudaFreeHost) for callers without a CUDA toolchain. */
void *HVQM4HostAlloc(size_t bytes);
void HVQM4HostFree(void *p);
/* Page-locks and maps memory the application already owns (e.g. the file images the bitstreams
   live in).  With HVQM4BatchSetEntropyMode(batch, 1), a step whose pictures all lie in registered
   ranges is fetched by the GPU itself (a gather kernel over PCIe): no host thread copies picture
   bytes, the host's share of a step is one 16-byte descriptor per picture.  Ranges may share pages
   (separately allocated buffers): a page stays registered until its last range is unregistered.
   It frees host cores, it is not faster (measured on B200, 1 024 streams, frames read back: dense
   97.6 k vs 105 k frames/s with 16 host threads, 95 k vs 92 k with 2; realistic 122 k vs 121 k and
   109 k vs 119 k), so nothing registers memory implicitly.
   LIFETIME: in this mode the GPU reads the picture bytes AFTER HVQM4BatchDecode has returned
   (frames[i] is otherwise only read during the call): they must stay unmodified until
   HVQM4BatchSync returns, or until the step four submissions later has been submitted (the ring of
   staging arenas).  HVQM4HostUnregister waits for the device to go idle first; unregister before
   the memory is freed.  Returns HVQM4_OK or error bits.
   PAGES: registration is by page, so other heap blocks of the application may share the first or last page of a
   range and thereby become partly page-locked.  CUDA refuses a copy whose host side straddles page-locked and pageable
   memory; the library's own read-backs and uploads (HVQM4BatchReadFrame*, the SDK entry points with host frames,
   HVQM4ConvertRGB) take such a frame over an internal page-locked bounce buffer instead of failing.  Page-aligned,
   page-padded ranges avoid the detour. */
int HVQM4HostRegister(void *ptr, size_t bytes);
int HVQM4HostUnregister(void *ptr);

/* ---------------------------------------------------------------- container helper
 * Minimal .h4m walker (header h4m:2175-2247, GOP blocks h4m:2429-2438, frame records
 * h4m:2456-2458) so that harnesses need not re-implement it.  Not needed by the decoder.
 */
typedef struct HVQM4FileInfo
{
    int32_t version;       /* 13 or 15 */
    int32_t width, height;
    int32_t h_samp, v_samp;
    int32_t n_gops, n_video_frames;
    int32_t usec_per_frame;
    /* audio track (header bytes 0x20, 0x3C-0x43, h4m:2196,2206-2210) */
    int32_t n_audio_frames;
    int32_t audio_channels, audio_bits, audio_format, audio_sample_rate;
} HVQM4FileInfo;

typedef struct HVQM4FrameRef
{
    uint32_t offset;       /* of the picture header inside the file image (record + 4) */
    uint32_t bytes;        /* record size - 4 */
    uint16_t frame_type;   /* 0x10 / 0x20 / 0x30 */
    uint16_t gop;
    uint32_t disp_id;      /* display index inside the GOP (h4m:2085) */
} HVQM4FrameRef;

/* Returns the number of video frames (<0 on a malformed container); fills at most
   max_frames entries. */
int HVQM4ParseFile(const uint8_t *data, size_t len, HVQM4FileInfo *info, HVQM4FrameRef *frames, int max_frames);

typedef struct HVQM4AudioRef
{
    uint32_t offset;       /* of the audio record payload (BE32 sample count, then data; h4m:2486-2488) */
    uint32_t bytes;        /* record size */
    uint16_t gop;
    uint16_t first;        /* 1: first audio record of its GOP block (state reset + seed, h4m:2446-2452) */
    uint32_t samples;      /* the payload's sample count */
} HVQM4AudioRef;

/* Same walk for the audio records: returns their number (<0 on a malformed container). */
int HVQM4ParseFileAudio(const uint8_t *data, size_t len, HVQM4AudioRef *refs, int max_refs);

/* ---------------------------------------------------------------- audio track
 * The IMA-ADPCM decoder of the reference (decode_audio, h4m:185-258; upstream keeps its call
 * disabled, h4m:2486-2507), batched over independent streams on the GPU (one thread per
 * stream: the recurrence is serial).  frames[i] = one audio record payload as HVQM4AudioRef
 * points at it, first[i] as in HVQM4AudioRef.  states[i] carries the predictor between the
 * records of a GOP block (zeroed by the library when first[i] is set).  pcm[i] receives
 * samples * channels int16 (interleaved, channel 0 first; native byte order), at most
 * pcm_capacity[i] samples; samples_out[i] (optional) = samples written.
 * Returns HVQM4_OK or error bits: HVQM4_ERR_TRUNCATED (payload shorter than its sample count
 * needs), HVQM4_ERR_ARGUMENT (seed step index > 88: the reference exits), HVQM4_ERR_OVERFLOW. */
#define HVQM4_AUDIO_MAX_CHANNELS 2
typedef struct HVQM4AudioState
{
    int16_t hist[HVQM4_AUDIO_MAX_CHANNELS];
    int8_t idx[HVQM4_AUDIO_MAX_CHANNELS];
    int8_t pad[2];
} HVQM4AudioState;

int HVQM4DecodeAudioBatch(int n, int channels, HVQM4AudioState *states, const int32_t *first, const uint8_t *const *frames,
                          const uint32_t *frame_bytes, int16_t *const *pcm, const uint32_t *pcm_capacity, uint32_t *samples_out);

/* ---------------------------------------------------------------- file player
 * The reference program as a library (main's container walk h4m:2427-2537 + decode_video
 * h4m:2078-2138): open a file image, then pull decoded pictures in file (= decode) order.  The
 * player owns the SeqObj, the work buffer and the three frame buffers and rotates them with the
 * reference's rule; pictures come back in host memory.  `data` must stay valid until Close.
 */
typedef struct HVQM4Player HVQM4Player;

HVQM4Player *HVQM4PlayerOpen(const uint8_t *data, size_t len);      /* NULL: malformed container, unsupported geometry or no device */
void HVQM4PlayerClose(HVQM4Player *p);
int HVQM4PlayerInfo(const HVQM4Player *p, HVQM4FileInfo *info);
/* Decodes the next video record.  *frame = the planar Y|U|V picture (valid until the next call),
   *display_index = video frames of the earlier GOP blocks + the record's disp_id (the number in
   the reference's output file name, h4m:2122), *frame_type = 0x10 / 0x20 / 0x30.
   Returns 1, 0 at the end of the file, < 0 on error. */
int HVQM4PlayerNextFrame(HVQM4Player *p, const uint8_t **frame, uint32_t *display_index, uint32_t *frame_type);
/* The reference's RGB conversion (dumpRGB, h4m:895-926) of the picture NextFrame returned last:
   width * height * 3 bytes.  Returns HVQM4_OK or error bits. */
int HVQM4PlayerFrameRGB(HVQM4Player *p, void *rgb);
/* Decodes the next audio record into pcm (interleaved int16, at most capacity samples per
   channel).  Returns the number of samples per channel, 0 at the end of the track, < 0 on error. */
int HVQM4PlayerNextAudio(HVQM4Player *p, int16_t *pcm, uint32_t capacity);
/* OR of the error bits raised since the last call (stream errors of the pictures, audio errors). */
uint32_t HVQM4PlayerErrors(HVQM4Player *p);

#ifdef __cplusplus
}
#endif
#endif
