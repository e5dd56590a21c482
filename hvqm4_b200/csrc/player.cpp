/*
 * player.cpp -- the reference PROGRAM as a library: container walk (main, h4m:2427-2537) and
 * the per-record decode with buffer rotation (decode_video, h4m:2078-2138), written on top of the
 * public C ABI of include/hvqm4.h only (SDK entry points, HVQM4ParseFile*, HVQM4ConvertRGB,
 * HVQM4DecodeAudioBatch).  "h4m:N" = /root/reference/h4m_audio_decode.c line N.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/hvqm4.h"

#define H4_API extern "C" __attribute__((visibility("default")))

struct HVQM4Player
{
    const uint8_t *data = nullptr;
    size_t len = 0;
    HVQM4FileInfo info{};
    std::vector<HVQM4FrameRef> frames;
    std::vector<HVQM4AudioRef> audio;
    std::vector<uint32_t> gop_start;         /* video frames before each GOP block (h4m:2471,2528) */
    size_t next_frame = 0, next_audio = 0;
    SeqObj seq{};
    VideoInfo vinfo{};
    void *work = nullptr;
    uint8_t *buf[3] = {nullptr, nullptr, nullptr};
    int past = 0, present = 1, future = 2;   /* h4m:2343-2349 */
    int shown = -1;                          /* buffer holding the picture NextFrame returned last */
    std::vector<uint8_t> record;             /* the picture + 3 bytes of slack (h4m:2080-2082) */
    HVQM4AudioState astate{};
    uint32_t errors = 0;
    bool seq_ready = false;
};

H4_API void HVQM4PlayerClose(HVQM4Player *p)
{
    if (!p) return;
    if (p->seq_ready) HVQM4ReleaseBuffer(&p->seq);
    free(p->work);
    for (uint8_t *b : p->buf) HVQM4HostFree(b);
    delete p;
}

H4_API HVQM4Player *HVQM4PlayerOpen(const uint8_t *data, size_t len)
{
    HVQM4Player *p = new HVQM4Player;
    p->data = data;
    p->len = len;
    const int nv = HVQM4ParseFile(data, len, &p->info, nullptr, 0);
    const int na = nv < 0 ? -1 : HVQM4ParseFileAudio(data, len, nullptr, 0);
    if (nv < 0 || na < 0)
    {
        delete p;
        return nullptr;
    }
    p->frames.resize((size_t)nv);
    p->audio.resize((size_t)na);
    if (nv) HVQM4ParseFile(data, len, &p->info, p->frames.data(), nv);
    if (na) HVQM4ParseFileAudio(data, len, p->audio.data(), na);
    p->gop_start.assign((size_t)p->info.n_gops + 1, 0);
    for (const HVQM4FrameRef &f : p->frames) ++p->gop_start[(size_t)f.gop + 1];
    for (size_t g = 1; g < p->gop_start.size(); ++g) p->gop_start[g] += p->gop_start[g - 1];

    /* the SDK call sequence of main(), h4m:2409-2419 */
    p->vinfo.hres = (uint16_t)p->info.width;
    p->vinfo.vres = (uint16_t)p->info.height;
    p->vinfo.h_samp = (uint8_t)p->info.h_samp;
    p->vinfo.v_samp = (uint8_t)p->info.v_samp;
    p->vinfo.video_mode = data[0x3A];
    HVQM4InitDecoder();
    HVQM4InitSeqObj(&p->seq, &p->vinfo);
    const uint32_t work_bytes = HVQM4BuffSize(&p->seq);
    p->work = calloc(1, work_bytes ? work_bytes : 1);
    const size_t frame_bytes = (size_t)p->info.width * p->info.height * 3 / 2;
    for (uint8_t *&b : p->buf)
    {
        b = static_cast<uint8_t *>(HVQM4HostAlloc(frame_bytes + 64));   /* pinned: the SDK calls copy to and from these */
        if (b) memset(b, 0, frame_bytes + 64);
    }
    if (!p->work || !p->buf[0] || !p->buf[1] || !p->buf[2])
    {
        HVQM4PlayerClose(p);
        return nullptr;
    }
    HVQM4SetBuffer(&p->seq, p->work);
    p->seq_ready = true;
    if (HVQM4SetVersion(&p->seq, p->info.version) != HVQM4_OK || (HVQM4GetLastError(&p->seq) & (HVQM4_ERR_NO_DEVICE | HVQM4_ERR_GEOMETRY)))
    {
        HVQM4PlayerClose(p);
        return nullptr;
    }
    return p;
}

H4_API int HVQM4PlayerInfo(const HVQM4Player *p, HVQM4FileInfo *info)
{
    if (!p || !info) return HVQM4_ERR_ARGUMENT;
    *info = p->info;
    return HVQM4_OK;
}

H4_API int HVQM4PlayerNextFrame(HVQM4Player *p, const uint8_t **frame, uint32_t *display_index, uint32_t *frame_type)
{
    if (!p || !frame) return -1;
    if (p->next_frame >= p->frames.size()) return 0;
    const HVQM4FrameRef &f = p->frames[p->next_frame++];
    const int t = f.frame_type;
    if (t != 0x30) { const int tmp = p->past; p->past = p->future; p->future = tmp; }         /* h4m:2087-2093 */
    p->record.resize((size_t)f.bytes + 8);
    memcpy(p->record.data(), p->data + f.offset, f.bytes);
    memset(p->record.data() + f.bytes, 0, 8);
    HVQM4SetFrameBytes(&p->seq, f.bytes);
    switch (t)
    {   /* h4m:2096-2104 */
    case 0x10: HVQM4DecodeIpic(&p->seq, p->record.data(), p->buf[p->present]); break;
    case 0x20: HVQM4DecodePpic(&p->seq, p->record.data(), p->buf[p->present], p->buf[p->past]); break;
    default: HVQM4DecodeBpic(&p->seq, p->record.data(), p->buf[p->present], p->buf[p->past], p->buf[p->future]); break;
    }
    const uint32_t e = HVQM4GetLastError(&p->seq);
    p->errors |= e;
    p->shown = p->present;
    *frame = p->buf[p->present];
    if (display_index) *display_index = p->gop_start[f.gop] + f.disp_id;                      /* h4m:2122 */
    if (frame_type) *frame_type = (uint32_t)t;
    if (t != 0x30) { const int tmp = p->present; p->present = p->future; p->future = tmp; }   /* h4m:2131-2137 */
    return (e & (HVQM4_ERR_NO_DEVICE | HVQM4_ERR_CUDA | HVQM4_ERR_NOMEM)) ? -2 : 1;
}

H4_API int HVQM4PlayerFrameRGB(HVQM4Player *p, void *rgb)
{
    if (!p || !rgb || p->shown < 0) return HVQM4_ERR_ARGUMENT;
    return HVQM4ConvertRGB(&p->seq, p->buf[p->shown], rgb);                                  /* h4m:2126 */
}

H4_API int HVQM4PlayerNextAudio(HVQM4Player *p, int16_t *pcm, uint32_t capacity)
{
    if (!p || !pcm) return -1;
    if (p->next_audio >= p->audio.size()) return 0;
    if (p->info.audio_channels < 1 || p->info.audio_channels > HVQM4_AUDIO_MAX_CHANNELS) return -3;
    const HVQM4AudioRef &a = p->audio[p->next_audio];
    const int32_t first = a.first;
    const uint8_t *payload = p->data + a.offset;
    const uint32_t bytes = a.bytes;
    uint32_t got = 0;
    const int rc = HVQM4DecodeAudioBatch(1, p->info.audio_channels, &p->astate, &first, &payload, &bytes, &pcm, &capacity, &got);
    if (rc & (HVQM4_ERR_OVERFLOW | HVQM4_ERR_NO_DEVICE | HVQM4_ERR_CUDA | HVQM4_ERR_NOMEM | (rc & HVQM4_ERR_ARGUMENT && !got ? HVQM4_ERR_ARGUMENT : 0))) return -2;
    ++p->next_audio;
    p->errors |= (uint32_t)rc;
    return (int)got;
}

H4_API uint32_t HVQM4PlayerErrors(HVQM4Player *p)
{
    if (!p) return 0;
    const uint32_t e = p->errors;
    p->errors = 0;
    return e;
}
