"""Independent float64 numpy model of the 802.11a/g OFDM PHY, written from the formulas of
IEEE 802.11-2012 clause 18 (not from the oracle, not from gr-ieee802-11): used to cross-check
the oracle's TX bit pipeline / waveform and its RX decisions.  It deliberately uses the
standard's own formulation where that differs from upstream's (forward interleaver
permutations 18-17/18-18, Gray tables, numpy FFT, full-traceback Viterbi)."""
import numpy as np

N_BPSC = [1, 1, 2, 2, 4, 4, 6, 6]
N_CBPS = [48, 48, 96, 96, 192, 192, 288, 288]
N_DBPS = [24, 36, 48, 72, 96, 144, 192, 216]
RATE_BITS = ["1101", "1111", "0101", "0111", "1001", "1011", "0001", "0011"]   # R1..R4
PUNCT = [(1, 1), (1, 1, 1, 0, 0, 1), (1, 1), (1, 1, 1, 0, 0, 1), (1, 1), (1, 1, 1, 0, 0, 1), (1, 1, 1, 0), (1, 1, 1, 0, 0, 1)]

LTS = np.array([1, 1, -1, -1, 1, 1, -1, 1, -1, 1, 1, 1, 1, 1, 1, -1, -1, 1, 1, -1, 1, -1, 1, 1, 1, 1, 0,
                1, -1, -1, 1, 1, -1, 1, -1, 1, -1, -1, -1, -1, -1, 1, 1, -1, -1, 1, -1, 1, -1, 1, 1, 1, 1], float)
STS_POS = {-24: 1, -20: -1, -16: 1, -12: -1, -8: -1, -4: 1, 4: -1, 8: -1, 12: 1, 16: 1, 20: 1, 24: 1}
DATA_SC = [k for k in range(-26, 27) if k not in (-21, -7, 0, 7, 21)]


def scrambler_seq(seed, n):
    s = [(seed >> (6 - i)) & 1 for i in range(7)]   # x7..x1
    out = []
    for _ in range(n):
        fb = s[0] ^ s[3]                            # x7 xor x4
        out.append(fb)
        s = s[1:] + [fb]
    return np.array(out, np.uint8)


PILOT_POLARITY = 1 - 2 * scrambler_seq(0x7f, 127).astype(int)


def conv_encode(bits):
    g0, g1 = 0o133, 0o171
    reg = 0
    out = []
    for b in bits:
        reg = ((reg << 1) | int(b)) & 0x7f          # newest bit at LSB
        # taps counted from the newest bit: g = 1 + D^2 + D^3 + D^5 + D^6 etc.
        a = sum(((reg >> d) & 1) for d in range(7) if (g0 >> (6 - d)) & 1) & 1
        c = sum(((reg >> d) & 1) for d in range(7) if (g1 >> (6 - d)) & 1) & 1
        out += [a, c]
    return np.array(out, np.uint8)


def puncture(coded, enc):
    pat = np.array(PUNCT[enc], bool)
    keep = np.resize(pat, coded.size)
    return coded[keep]


def interleave_perm(enc):
    """Standard 18-17 / 18-18: bit k of the encoder output goes to position j."""
    n, s = N_CBPS[enc], max(N_BPSC[enc] // 2, 1)
    k = np.arange(n)
    i = (n // 16) * (k % 16) + k // 16
    j = s * (i // s) + (i + n - (16 * i) // n) % s
    return j


def gray_axis(bits):
    table = {1: {(0,): -1, (1,): 1},
             2: {(0, 0): -3, (0, 1): -1, (1, 1): 1, (1, 0): 3},
             3: {(0, 0, 0): -7, (0, 0, 1): -5, (0, 1, 1): -3, (0, 1, 0): -1, (1, 1, 0): 1, (1, 1, 1): 3, (1, 0, 1): 5, (1, 0, 0): 7}}
    return table[len(bits)][tuple(int(b) for b in bits)]


def map_bits(bits, enc):
    nb = N_BPSC[enc]
    b = np.asarray(bits).reshape(-1, nb)
    if nb == 1:
        return (2.0 * b[:, 0] - 1).astype(complex)
    h = nb // 2
    norm = {1: np.sqrt(2), 2: np.sqrt(10), 3: np.sqrt(42)}[h]
    return np.array([(gray_axis(r[:h]) + 1j * gray_axis(r[h:])) / norm for r in b])


def signal_bits(enc, length):
    b = [int(c) for c in RATE_BITS[enc]] + [0] + [(length >> i) & 1 for i in range(12)]
    b.append(sum(b) & 1)
    return np.array(b + [0] * 6, np.uint8)


def data_bits(psdu, enc, seed):
    n_sym = -(-(16 + 8 * len(psdu) + 6) // N_DBPS[enc])
    n_data = n_sym * N_DBPS[enc]
    bits = np.zeros(n_data, np.uint8)
    bits[16:16 + 8 * len(psdu)] = np.unpackbits(np.frombuffer(bytes(psdu), np.uint8), bitorder="little")
    scr = bits ^ scrambler_seq(seed, n_data)
    scr[16 + 8 * len(psdu):16 + 8 * len(psdu) + 6] = 0
    return scr, n_sym


def ofdm_symbol(data48, pilot_pol):
    X = np.zeros(64, complex)                      # index = subcarrier mod 64
    for v, k in zip(data48, DATA_SC):
        X[k % 64] = v
    for k, sgn in zip((-21, -7, 7, 21), (1, 1, 1, -1)):
        X[k % 64] = pilot_pol * sgn
    return X


def time_symbol(X):
    return np.fft.ifft(X) * 64 / np.sqrt(52)


def tx_frame(psdu, enc, seed):
    """Returns (samples, per-symbol carrier indices as the mapper numbers them: bit k = b_k)."""
    scr, n_sym = data_bits(psdu, enc, seed)
    coded = puncture(conv_encode(scr), enc)
    perm = interleave_perm(enc)
    ncb, nb = N_CBPS[enc], N_BPSC[enc]
    syms_f = []
    S = np.zeros(64, complex)
    for k, s in STS_POS.items():
        S[k % 64] = s * np.sqrt(13 / 6) * (1 + 1j)
    L = np.zeros(64, complex)
    for i, k in enumerate(range(-26, 27)):
        L[k % 64] = LTS[i]
    sts_t, lts_t = time_symbol(S), time_symbol(L)
    parts = [np.tile(sts_t, 3)[:160], np.concatenate([lts_t[32:], lts_t, lts_t])]
    sig = np.zeros(48, np.uint8)
    sig[interleave_perm(0)] = conv_encode(signal_bits(enc, len(psdu)))
    t = time_symbol(ofdm_symbol(map_bits(sig, 0), PILOT_POLARITY[0]))
    parts.append(np.concatenate([t[48:], t]))
    idx = np.zeros((n_sym, 48), np.uint8)
    for n in range(n_sym):
        blk = np.zeros(ncb, np.uint8)
        blk[perm] = coded[n * ncb:(n + 1) * ncb]
        idx[n] = (blk.reshape(48, nb) << np.arange(nb)).sum(1)
        t = time_symbol(ofdm_symbol(map_bits(blk, enc), PILOT_POLARITY[(n + 1) % 127]))
        parts.append(np.concatenate([t[48:], t]))
    # time windowing as the GNU Radio cyclic prefixer does it (rolloff 2): symbol boundaries are
    # the average of the incoming symbol's first CP sample and the periodic extension of the previous one
    out = np.concatenate(parts + [np.zeros(1, complex)])
    starts = list(range(0, 80 * (5 + n_sym) + 1, 80))
    # periodic extension of each 80-sample symbol = its sample 16 (x[0] of the 64-point body)
    bodies = np.concatenate(parts).reshape(-1, 80)
    for s_i, st in enumerate(starts):
        prev = bodies[s_i - 1][16] if s_i > 0 else 0.0
        cur = bodies[s_i][0] if s_i < len(bodies) else 0.0
        out[st] = 0.5 * cur + 0.5 * prev
    return out, idx


# ---------------------------------------------------------------- receive side (genie timing)
def viterbi_full(sym):
    """Hard-decision ML decoder, erasures = 2, full traceback from the best end state."""
    n = len(sym) // 2
    nxt = np.zeros((64, 2), int)
    outb = np.zeros((64, 2, 2), int)
    for s in range(64):
        for b in range(2):
            reg = ((s << 1) | b) & 0x7f
            nxt[s, b] = reg & 0x3f
            outb[s, b, 0] = bin(reg & 0o155).count("1") & 1
            outb[s, b, 1] = bin(reg & 0o117).count("1") & 1
    metric = np.zeros(64)
    back = np.zeros((n, 64), np.int8)
    prev_of = [[(s2 >> 1), (s2 >> 1) | 32] for s2 in range(64)]
    for t in range(n):
        r0, r1 = sym[2 * t], sym[2 * t + 1]
        new = np.full(64, -1e9)
        for s2 in range(64):
            b = s2 & 1
            for which, s in enumerate(prev_of[s2]):
                m = metric[s]
                if r0 != 2:
                    m += outb[s, b, 0] == r0
                if r1 != 2:
                    m += outb[s, b, 1] == r1
                if m > new[s2]:
                    new[s2] = m
                    back[t, s2] = which
        metric = new
    s = int(np.argmax(metric))
    bits = np.zeros(n, np.uint8)
    for t in range(n - 1, -1, -1):
        bits[t] = s & 1
        s = prev_of[s][back[t, s]]
    return bits


def rx_frame(samples, frame_start, enc=None):
    """Genie-timed LS receiver: returns (enc, length, psdu bytes or None)."""
    x = np.asarray(samples, complex)
    l1 = np.fft.fft(x[frame_start + 192:frame_start + 256]) * np.sqrt(52) / 64
    l2 = np.fft.fft(x[frame_start + 256:frame_start + 320]) * np.sqrt(52) / 64
    Lf = np.zeros(64)
    for i, k in enumerate(range(-26, 27)):
        Lf[k % 64] = LTS[i]
    used = Lf != 0
    H = np.ones(64, complex)
    H[used] = (l1[used] + l2[used]) / 2 / Lf[used]

    def demod(n_sym_index, enc_):
        st = frame_start + 320 + 80 * n_sym_index + 16
        Y = np.fft.fft(x[st:st + 64]) * np.sqrt(52) / 64 / H
        pil = sum(Y[k % 64] * sgn for k, sgn in zip((-21, -7, 7, 21), (1, 1, 1, -1))) * PILOT_POLARITY[n_sym_index % 127]
        Y = Y * np.exp(-1j * np.angle(pil))
        d = np.array([Y[k % 64] for k in DATA_SC])
        nb = N_BPSC[enc_]
        if nb == 1:
            return (d.real > 0).astype(np.uint8).reshape(48, 1)
        h = nb // 2
        norm = {1: np.sqrt(2), 2: np.sqrt(10), 3: np.sqrt(42)}[h]
        levels = {1: [-1, 1], 2: [-3, -1, 1, 3], 3: [-7, -5, -3, -1, 1, 3, 5, 7]}[h]
        inv = {gray_axis(b): b for b in [tuple((v >> (h - 1 - i)) & 1 for i in range(h)) for v in range(1 << h)]}
        out = np.zeros((48, nb), np.uint8)
        for c in range(48):
            for ax, val in enumerate((d[c].real * norm, d[c].imag * norm)):
                lv = min(levels, key=lambda q: abs(q - val))
                out[c, ax * h:(ax + 1) * h] = inv[lv]
        return out

    sig = demod(0, 0).reshape(-1)
    dec = viterbi_full(sig[interleave_perm(0)])
    rate = "".join(str(b) for b in dec[:4])
    if rate not in RATE_BITS or (dec[:17].sum() & 1) != dec[17]:
        return None, None, None
    enc_ = RATE_BITS.index(rate)
    length = int(sum(int(dec[5 + i]) << i for i in range(12)))
    n_sym = -(-(16 + 8 * length + 6) // N_DBPS[enc_])
    perm = interleave_perm(enc_)
    coded = np.concatenate([demod(1 + n, enc_).reshape(-1)[perm] for n in range(n_sym)])
    pat = np.resize(np.array(PUNCT[enc_], bool), 2 * n_sym * N_DBPS[enc_])
    dep = np.full(pat.size, 2, np.uint8)
    dep[pat] = coded
    bits = viterbi_full(dep)
    seq_state = bits[:7]
    seed = int(sum(int(b) << (6 - i) for i, b in enumerate(seq_state)))
    # the first 7 scrambled SERVICE bits are the scrambler output itself: regenerate the sequence
    for cand in range(1, 128):
        if np.array_equal(scrambler_seq(cand, 7), seq_state):
            seed = cand
            break
    plain = bits ^ scrambler_seq(seed, bits.size)
    return enc_, length, np.packbits(plain[16:16 + 8 * length], bitorder="little").tobytes()
