/*
 * Sanitizer harness for the host serial stage (hvqm4_b200/csrc/entropy.c), test infrastructure only.
 * Built by tests/test_host_stage_sanitized.py with -fsanitize=address,undefined.
 *
 *   entropy_san <file.h4m> <list.txt> <width> <height> <version15> <rounds> <seed> [split]
 *
 * list.txt: one "offset bytes type" line per picture, in decode order (the container walk is
 * the test's job).  Round 0 parses the pictures as they are; every further round parses a damaged
 * copy of each picture: random byte flips, corrupted header / section table, or truncation.  Each
 * picture is handed over in a heap block of EXACTLY its readable size, and the symbol buffer is a
 * heap block of EXACTLY the size h4e_parse_begin asked for, so that any read past the picture and
 * any write past the blob is an AddressSanitizer report (h4m:2080-2082 lets the reference read
 * three bytes past the record; this stage claims to need none).
 * Every parsed picture is then RECONSTRUCTED by tests/emul/recon_emul.cpp -- the block functions and the
 * addressing of the CUDA kernels (recon_core.h) run serially -- into frame surfaces that are heap blocks of
 * exactly frame bytes + 64 (the tail slack the library gives its device surfaces): a vector, a window or a
 * record of a damaged picture that reached outside a surface or the blob would be a report here, where
 * compute-sanitizer cannot be run on the GPU pool.
 * split = 1 selects the pass structure of the GPU build (h4e_seq_set_split_schedule): the device
 * code is this same file, so this is also the closest a CPU sanitizer gets to the GPU parser.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "entropy.h"

/* tests/emul/recon_emul.cpp */
int emul_recon_picture(const uint8_t *blob, uint8_t *present, const uint8_t *past, const uint8_t *future);

static uint64_t rng_state;
static uint32_t rnd(void)
{
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 16);
}

int main(int argc, char **argv)
{
    if (argc < 8) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    fseek(f, 0, SEEK_END);
    const long flen = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *file = malloc((size_t)flen);
    if (fread(file, 1, (size_t)flen, f) != (size_t)flen) return 2;
    fclose(f);
    const int width = atoi(argv[3]), height = atoi(argv[4]), v15 = atoi(argv[5]), rounds = atoi(argv[6]);
    rng_state = 0x9E3779B97F4A7C15ull ^ (uint64_t)atoll(argv[7]);
    const int split = argc > 8 ? atoi(argv[8]) : 0;

    long off[4096], len[4096];
    int type[4096], n = 0;
    FILE *l = fopen(argv[2], "r");
    if (!l) return 2;
    while (n < 4096 && fscanf(l, "%ld %ld %d", &off[n], &len[n], &type[n]) == 3) ++n;
    fclose(l);

    unsigned long parsed = 0, flagged = 0, refused = 0, painted = 0;
    uint32_t all_bits = 0;
    const size_t surf_bytes = (size_t)width * height * 3 / 2 + 64;
    for (int round = 0; round < rounds; ++round)
    {
        H4Seq *s = h4e_seq_create(width, height, 2, 2, v15);
        if (!s) return 3;
        h4e_seq_set_split_schedule(s, split);
        uint8_t *surf[3];
        for (int k = 0; k < 3; ++k) surf[k] = calloc(1, surf_bytes);
        int past = 0, present = 1, future = 2;                            /* rotation of decode_video, h4m:2087-2093, 2131-2137 */
        for (int i = 0; i < n; ++i)
        {
            if (type[i] != 0x30) { const int t = past; past = future; future = t; }
            if (off[i] < 0 || len[i] <= 0 || off[i] + len[i] > flen) return 2;
            size_t bytes = (size_t)len[i];
            const int mode = round == 0 ? 0 : 1 + (int)(rnd() % 3);
            if (mode == 3) bytes = 1 + rnd() % bytes;                     /* truncated record */
            uint8_t *pic = malloc(bytes);                                 /* exact size: overreads are reports */
            memcpy(pic, file + off[i], bytes);
            if (mode == 1)
                for (int k = 0, flips = 1 + (int)(rnd() % 24); k < flips; ++k) pic[rnd() % bytes] ^= (uint8_t)(1u << (rnd() & 7));
            if (mode == 2)
                for (int k = 0, flips = 1 + (int)(rnd() % 6); k < flips; ++k) pic[rnd() % (bytes < 76 ? bytes : 76)] = (uint8_t)rnd();
            const size_t need = h4e_parse_begin(s, type[i], pic, bytes);
            if (need)
            {
                uint8_t *blob = malloc(need);                             /* exact size: overruns are reports */
                const uint32_t bits = h4e_parse_finish(s, blob);
                all_bits |= bits;
                flagged += bits != 0;
                ++parsed;
                const int rc = emul_recon_picture(blob, surf[present], surf[past], type[i] == 0x20 ? surf[present] : surf[future]);
                if (rc != 0 && bits == 0) return 4;                       /* an intact picture must reconstruct */
                painted += rc == 0;
                free(blob);
            }
            else
                ++refused;
            free(pic);
            if (type[i] != 0x30) { const int t = present; present = future; future = t; }
        }
        for (int k = 0; k < 3; ++k) free(surf[k]);
        h4e_seq_destroy(s);
    }
    free(file);
    printf("%lu parsed %lu flagged %lu refused %lu reconstructed bits 0x%x\n", parsed, flagged, refused, painted, all_bits);
    return 0;
}
