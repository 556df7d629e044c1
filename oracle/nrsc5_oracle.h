/* nrsc5_oracle — plain-C CPU restatement of the NRSC-5 FM physical-layer
 * receive chain of theori-io/nrsc5 (reference @ a5c0972).
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA product in
 * nrsc5_b200/csrc; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product never links,
 * imports or calls anything under oracle/.
 *
 * Parity pin: tests/test_oracle.py checks this restatement against the
 * UNMODIFIED reference compiled into oracle/_ref/libnrsc5_ref.so (soft bits
 * bit-identical, L1 PDUs identical, events identical) on support/sample.xz and
 * on synthetic captures, and against golden digests committed under
 * tests/golden/.
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* log record types: identical to oracle/reftap.c so one parser serves both */
enum {
    ORC_REC_FRAME = 1,     /* u32 lc, u32 nbits, bits packed MSB-first  */
    ORC_REC_PIDS = 2,      /* 10 bytes + u8 CRC-12 verdict (pids.c:52-86) */
    ORC_REC_SYNC = 3,      /* f32 freq_offset, i32 psmi, pli, hppi, aabi, rdbi */
    ORC_REC_LOST_SYNC = 4,
    ORC_REC_MER = 5,       /* f32 lower, f32 upper                      */
    ORC_REC_BER = 6,       /* f32 cber                                  */
    ORC_REC_SOFT_PM = 8,   /* u32 bc, 23040 int8                        */
    ORC_REC_BLOCK = 9,     /* i32 state_in, i32 samperr, f32 angle, f32 ph_re, f32 ph_im, i32 cfo, i64 start */
};

/* the L2 -> L3 calls of frame.c, same layout as oracle/reftap_l2.c */
enum {
    ORC_REC_L2_SERVICE = 16,   /* i32 x8: program, access, type, codec_mode, blend_control, gain, common_delay, latency */
    ORC_REC_L2_ALIGN = 17,     /* u32 program, stream_id, offset                                                        */
    ORC_REC_L2_AAS = 18,       /* the bytes handed to output_aas_push                                                   */
    ORC_REC_L2_PACKET = 19,    /* u32 program, stream_id, seq, shape, flags, size; then the packet bytes                */
};

typedef struct orc orc_t;

orc_t *orc_new(void);
void orc_free(orc_t *o);
void orc_want_soft(orc_t *o, int on);
void orc_want_blocks(orc_t *o, int on);
/* mirrors input_push_cu8 (reference src/input.c:96); nbytes % 4 == 0 */
void orc_push_cu8(orc_t *o, const uint8_t *buf, size_t nbytes);
/* mirrors input_push_cs16 (reference src/input.c:119); nvalues % 2 == 0 */
void orc_push_cs16(orc_t *o, const int16_t *buf, size_t nvalues);
size_t orc_log_size(const orc_t *o);
const uint8_t *orc_log_data(const orc_t *o);
void orc_log_clear(orc_t *o);

/* ---- L2 framing (oracle/nrsc5_oracle_l2.c): frame_push / frame_process, reference src/frame.c:130-714 ---- */
typedef struct orc_l2 orc_l2_t;
orc_l2_t *orc_l2_new(void);
void orc_l2_free(orc_l2_t *o);
void orc_l2_reset(orc_l2_t *o);                                   /* frame_reset */
void orc_l2_push(orc_l2_t *o, const uint8_t *packed, unsigned nbits, unsigned lc);   /* frame_push */
/* {u32 lc, u32 nbits, packed bits padded to 4} back to back; nbits == 0 = frame_reset */
int orc_l2_frames(orc_l2_t *o, const uint8_t *frames, size_t nbytes);
size_t orc_l2_log_size(const orc_l2_t *o);
const uint8_t *orc_l2_log_data(const orc_l2_t *o);
void orc_l2_log_clear(orc_l2_t *o);
unsigned orc_l2_lost(const orc_l2_t *o);                          /* sync-loss predicate count (frame.c:535-540) */

/* ---- AM (hybrid MA1), oracle/nrsc5_oracle_am.c: cs16 at 46 511.72 S/s in, same record stream out ---- */
/* CRC-12 verdict of a PIDS frame (80 bits packed MSB-first), reference src/pids.c:52-86,1032-1050 */
int orc_pids_crc12_ok(const uint8_t *pk);

typedef struct orc_am orc_am_t;
orc_am_t *orc_am_new(void);
void orc_am_free(orc_am_t *o);
/* mirrors input_push_cs16 in AM mode (reference src/input.c:119); nvalues % 2 == 0 */
void orc_am_push_cs16(orc_am_t *o, const int16_t *buf, size_t nvalues);
/* cu8 at 1 488 375 S/s (input_push_cu8 in AM mode: /32 through five halfband stages); nbytes % 4 == 0 */
void orc_am_push_cu8(orc_am_t *o, const uint8_t *buf, size_t nbytes);
/* that decimator alone, from a zero state: returns the number of cs16 complex samples written (nbytes / 64) */
size_t orc_am_decimate(const uint8_t *cu8, size_t nbytes, int16_t *out);
size_t orc_am_log_size(const orc_am_t *o);
const uint8_t *orc_am_log_data(const orc_am_t *o);

/* ---- stage-level functions (pure; for kernel-by-kernel parity tests) ---- */
/* cu8 -> Q15 -> halfband /2 from a zero history; n_out = npairs */
void orc_halfband_fm(const uint8_t *cu8, size_t npairs, int16_t *out_ri);
/* tail-biting Viterbi, n=3; k = 7 or 9; in: 3*len int8; out: len bits (one per byte) */
void orc_viterbi(const int8_t *in, uint8_t *out, int k, int len, unsigned g0, unsigned g1, unsigned g2);
/* interleaver I + depuncture for P1: pm[16*23040] -> out[438528] */
void orc_deinterleave_p1(const int8_t *pm, int8_t *out);
/* interleaver II + depuncture for PIDS of block bc: pm -> out[240] */
void orc_deinterleave_pids(const int8_t *pm, unsigned bc, int8_t *out);
void orc_descramble(uint8_t *bits, unsigned len);
/* channel bit errors by re-encoding (rate 2/5 FM) */
int orc_bit_errors_fm(const int8_t *coded, const uint8_t *decoded, int len);
/* RS(255,247) decode in place; returns corrections or -1 */
int orc_rs_decode(uint8_t *block255);
/* L2 header fix on the first 96 PDU bytes; returns 1 ok / 0 fail */
int orc_fix_header(uint8_t *buf96);
/* L2 sync-loss predicate on a descrambled P1 frame (146176 bits, one per
 * byte): returns 1 when the reference would drop to SYNC_STATE_NONE */
int orc_p1_sync_lost(const uint8_t *bits, uint32_t *pci_out);

#ifdef __cplusplus
}
#endif
