/*
 * c_abi_demo.c -- the drop-in boundary from plain C: no Python, no CUDA headers, only include/wifi_b200.h.
 *
 * What gnu_radio/IRS_tranceiver.py does around the hier block (mac -> wifi_phy_hier TX (:178-184,313) -> x0.6 (:295) ->
 * pad 100/1000 (:277) -> x sqrt(10^(snr/10)) (:294) -> channel_model (:282-288) -> RX chain (:268-273) -> "Extract Pics"
 * IRS_tranceiver_epy_block_2.py:31-38), with every PHY step inside libwifi_b200.so on the GPU.
 *
 *   gcc -O2 -I include examples/c_abi_demo.c -o c_abi_demo -L gnuradio-wifi-imagetransfer_b200 -lwifi_b200 \
 *       -Wl,-rpath,$PWD/gnuradio-wifi-imagetransfer_b200 -lm
 *   ./c_abi_demo [n_frames] [encoding 0..7] [snr slider dB]
 *
 * Exit code 0: every payload came back bit-exact.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "wifi_b200.h"

#define CHECK(call)                                                                                   \
    do {                                                                                              \
        long long rc_ = (long long)(call);                                                            \
        if (rc_ < 0) {                                                                                \
            fprintf(stderr, "%s -> %s (%s)\n", #call, wifi_b200_strerror((int)rc_), h ? wifi_b200_last_error(h) : ""); \
            return 2;                                                                                 \
        }                                                                                             \
    } while (0)

int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 16;
    const int enc = argc > 2 ? atoi(argv[2]) : WIFI_QPSK_3_4;      /* the loopback flowgraph's default (IRS_tranceiver.py:91) */
    const double snr = argc > 3 ? atof(argv[3]) : 27.0;
    const int payload_len = 268;                                    /* one pickled 10x10x1 uint8 patch (upload_image_udp.py:29-32) */
    const uint8_t src[6] = {0x23, 0x23, 0x23, 0x23, 0x23, 0x23}, dst[6] = {0x42, 0x42, 0x42, 0x42, 0x42, 0x42},
                  bss[6] = {0xff, 0xff, 0xff, 0xff, 0xff, 0xff};
    wifi_b200_t *h = NULL;
    if (n < 1 || enc < 0 || enc > 7) return 2;
    if (wifi_b200_device_count() < 1) {
        fprintf(stderr, "no sm_100 GPU: the library has no CPU path\n");
        return 3;
    }
    const int psdu_len = payload_len + 28;
    const int flen = wifi_b200_frame_samples(enc, psdu_len);
    const size_t burst = 100 + (size_t)flen + 1000;                 /* foo.packet_pad2(100, 1000) */
    wifi_b200_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.bandwidth = 20e6; cfg.frequency = 5.89e9; cfg.sensitivity = 0.56;
    cfg.chan_est = WIFI_EQ_LS; cfg.encoding = enc; cfg.min_plateau = 2;
    cfg.max_samples = (int64_t)(n * burst + 1024); cfg.max_frames = n + 64;
    CHECK(wifi_b200_create(&cfg, &h));

    /* mac 'app in' -> 'phy out' */
    uint8_t *payloads = malloc((size_t)n * payload_len), *psdus = malloc((size_t)n * psdu_len);
    uint32_t *off = malloc(sizeof(uint32_t) * n), *len = malloc(sizeof(uint32_t) * n);
    srand(1);
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < payload_len; ++k) payloads[(size_t)i * payload_len + k] = (uint8_t)rand();
        CHECK(wifi_b200_mac_frame(payloads + (size_t)i * payload_len, payload_len, i, src, dst, bss, psdus + (size_t)i * psdu_len));
        off[i] = (uint32_t)(i * psdu_len);
        len[i] = (uint32_t)psdu_len;
    }
    /* mac_in -> samp_out */
    float *tx = malloc(sizeof(float) * 2 * (size_t)n * flen);
    uint64_t *boff = malloc(sizeof(uint64_t) * (n + 1));
    CHECK(wifi_b200_tx(h, psdus, off, len, NULL, NULL, n, tx, (size_t)n * flen, boff));
    /* x0.6, pad, x sqrt(10^(snr/10)), channel_model(noise_voltage 1): one descriptor per padded burst */
    float *rx = malloc(sizeof(float) * 2 * n * burst);
    wifi_b200_chan_seg *seg = calloc(n, sizeof *seg);
    for (int i = 0; i < n; ++i) {
        seg[i].in_off = (int64_t)boff[i];                           /* reads outside the burst return 0: the padding */
        seg[i].in_len = flen;
        seg[i].out_off = (int64_t)(i * burst);
        seg[i].n = (int64_t)burst;
        seg[i].n0 = seg[i].out_off;                                 /* Philox counter = absolute sample index: one noise stream */
        seg[i].gain = (float)(0.6 * sqrt(pow(10.0, snr / 10.0)));
        seg[i].noise_sigma = (float)sqrt(2.0);                      /* noise_voltage 1 = unit variance per I/Q component */
        seg[i].n_taps = 1;
        seg[i].delay[0] = 100;                                      /* 100 zeros in front, 1000 behind */
        seg[i].tap_re[0] = 1.f;
        seg[i].seed = 0;
    }
    CHECK(wifi_b200_channel(h, tx, (int64_t)n * flen, rx, (int64_t)(n * burst), seg, n));
    /* samp_in -> mac_out */
    uint64_t link_off[2] = {0, (uint64_t)(n * burst)};
    CHECK(wifi_b200_rx_batch(h, rx, link_off, 1, 1));
    int64_t nf = 0, nrows = 0, npdu = 0, store = 0;
    CHECK(wifi_b200_rx_counts(h, &nf, &nrows, &npdu, &store));
    wifi_b200_frame *fr = malloc(sizeof *fr * (size_t)(nf ? nf : 1));
    uint8_t *st = malloc((size_t)(store ? store : 1));
    if (nf) CHECK(wifi_b200_rx_frames(h, fr, nf));
    if (store) CHECK(wifi_b200_rx_psdus(h, st, (size_t)store));
    int good = 0;
    for (int64_t i = 0; i < nf; ++i) {
        if (!fr[i].crc_ok || fr[i].length != psdu_len) continue;
        const uint8_t *mpdu = st + fr[i].psdu_off;                  /* PSDU incl. FCS; mac_out carries length - 4 bytes */
        const int seq = (mpdu[22] | (mpdu[23] << 8)) >> 4;          /* which datagram this is */
        /* "Extract Pics" forwards data[24:][4:]; compare the whole payload */
        if (seq < n && memcmp(mpdu + 24, payloads + (size_t)seq * payload_len, payload_len) == 0) ++good;
    }
    wifi_b200_stats s;
    CHECK(wifi_b200_get_stats(h, &s));
    printf("%d frames sent (%s-like MCS %d, slider snr %.1f dB), %lld triggers, %lld PDUs, %d payloads bit-exact, %lld samples\n", n,
           "IRS_tranceiver", enc, snr, (long long)nf, (long long)npdu, good, (long long)s.samples);
    wifi_b200_destroy(h);
    return good == n ? 0 : 1;
}
