/* reftap_l2 — test-infrastructure taps at the reference's L2 -> L3 calls
 * (SURVEY §8 f1): the UNMODIFIED reference frame.c (src/frame.c:516-714) is
 * observed, not restated.  Link-time --wrap taps record, in call order,
 *   output_align()                reference src/frame.c:606   (src/output.c:31)
 *   output_push()                 reference src/frame.c:635   (src/output.c:47)
 *   output_aas_push()             reference src/frame.c:365   (src/output.c)
 *   nrsc5_report_audio_service()  reference src/frame.c:590   (src/nrsc5.c)
 * into the same log reftap.c writes (so they interleave with REC_FRAME etc.).
 * reftap_l2_frames() feeds a sequence of L1 PDUs straight into the reference's
 * frame_push() on a fresh handle, so L2 can be pinned on hand-made PDUs
 * without modulating a capture.
 * Compiled with the reference's own headers (it needs the layout of nrsc5_t).
 * Nothing here is part of the product.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "private.h"

enum {
    REC_L2_SERVICE = 16,   /* i32 x8: program, access, type, codec_mode, blend_control, gain, common_delay, latency */
    REC_L2_ALIGN = 17,     /* u32 program, stream_id, offset                                                        */
    REC_L2_AAS = 18,       /* bytes handed to output_aas_push                                                       */
    REC_L2_PACKET = 19,    /* u32 program, stream_id, seq, shape, flags, size; then the packet bytes                */
};

static int g_want_l2;          /* the L2 -> L3 records are only written on request (reftap_want_l2, reftap_l2_frames) */
void reftap_want_l2(int on) { g_want_l2 = on; }

void reftap_log_put(uint32_t type, const void *a, size_t alen, const void *b, size_t blen);
void reftap_set_logging(int on);

void __real_output_align(output_t *st, unsigned int program, unsigned int stream_id, unsigned int offset);
void __real_output_push(output_t *st, const packet_ref_t *ref);
void __real_output_aas_push(output_t *st, uint8_t *psd, unsigned int len);
void __real_nrsc5_report_audio_service(nrsc5_t *st, unsigned int program, unsigned int access, unsigned int type,
                                       unsigned int codec_mode, unsigned int blend_control, int digital_audio_gain,
                                       unsigned int common_delay, unsigned int latency);
void __real_frame_push(frame_t *st, uint8_t *bits, size_t length, logical_channel_t lc);
void __wrap_frame_push(frame_t *st, uint8_t *bits, size_t length, logical_channel_t lc);

void __wrap_output_align(output_t *st, unsigned int program, unsigned int stream_id, unsigned int offset)
{
    uint32_t p[3] = { program, stream_id, offset };
    if (g_want_l2) reftap_log_put(REC_L2_ALIGN, p, sizeof(p), NULL, 0);
    __real_output_align(st, program, stream_id, offset);
}

void __wrap_output_push(output_t *st, const packet_ref_t *ref)
{
    uint32_t p[6] = { ref->program, ref->stream_id, ref->seq, ref->shape, ref->flags, ref->size };
    if (g_want_l2) reftap_log_put(REC_L2_PACKET, p, sizeof(p), ref->data, ref->size);
    __real_output_push(st, ref);
}

/* reftap_l2_frames() feeds generated PDUs whose AAS payloads are random bytes; the L3 parsers behind
 * output_aas_push (ID3, SIG, LOT: outside the L2 row) are not meant to see those, so the call is recorded there
 * and not forwarded.  Captures decoded through the public API (reftap_decode) are forwarded as always. */
static int g_isolate_l3;

void __wrap_output_aas_push(output_t *st, uint8_t *psd, unsigned int len)
{
    if (g_want_l2) reftap_log_put(REC_L2_AAS, psd, len, NULL, 0);
    if (!g_isolate_l3)
        __real_output_aas_push(st, psd, len);
}

void __wrap_nrsc5_report_audio_service(nrsc5_t *st, unsigned int program, unsigned int access, unsigned int type,
                                       unsigned int codec_mode, unsigned int blend_control, int digital_audio_gain,
                                       unsigned int common_delay, unsigned int latency)
{
    int32_t p[8] = { (int32_t)program, (int32_t)access, (int32_t)type, (int32_t)codec_mode, (int32_t)blend_control,
                     digital_audio_gain, (int32_t)common_delay, (int32_t)latency };
    if (g_want_l2) reftap_log_put(REC_L2_SERVICE, p, sizeof(p), NULL, 0);
    __real_nrsc5_report_audio_service(st, program, access, type, codec_mode, blend_control, digital_audio_gain,
                                      common_delay, latency);
}

/* frames: nframes entries of {u32 lc, u32 nbits, packed bits MSB-first padded to 4 bytes}, back to back; an entry
 * with nbits == 0 stands for frame_reset() (what entering fine sync does, reference src/sync.c:405-409).
 * mode: NRSC5_MODE_FM / NRSC5_MODE_AM (the lost-sync feedback of AM P1 frames depends on the PDU length only). */
int reftap_l2_frames(const uint8_t *frames, size_t nbytes, int mode)
{
    nrsc5_t *st = NULL;
    if (nrsc5_open_pipe(&st) != 0)
        return -1;
    nrsc5_set_mode(st, mode);
    reftap_set_logging(1);
    g_isolate_l3 = 1;
    const int want_before = g_want_l2;
    g_want_l2 = 1;
    size_t off = 0;
    uint8_t *bits = (uint8_t *)malloc(P1_FRAME_LEN_FM);
    while (off + 8 <= nbytes) {
        uint32_t hdr[2];
        memcpy(hdr, frames + off, 8);
        off += 8;
        if (hdr[1] == 0) {
            frame_reset(&st->input.frame);
            continue;
        }
        size_t nb = (hdr[1] + 7) / 8;
        for (uint32_t i = 0; i < hdr[1]; i++)
            bits[i] = (frames[off + (i >> 3)] >> (7 - (i & 7))) & 1;
        off += (nb + 3) & ~(size_t)3;
        __wrap_frame_push(&st->input.frame, bits, hdr[1], (logical_channel_t)hdr[0]);
    }
    free(bits);
    g_isolate_l3 = 0;
    g_want_l2 = want_before;
    reftap_set_logging(0);
    nrsc5_close(st);
    return 0;
}
