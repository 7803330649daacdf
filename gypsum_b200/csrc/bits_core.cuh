// Pseudosymbol -> navigation bit integration (reference gypsum/navigation_bit_intergrator.py:105-288) as a per-channel
// state machine over the +-1 stream the tracking kernel leaves in device memory: 20 symbols -> one 50 bps bit, with the
// reference's bit-phase search, health-driven resynchronisation and queue/cursor bookkeeping (including Python's
// negative-index slice semantics for a cursor that went negative after a phase change, :266-269).
// Host/device code: the device runs it on one lane per channel (bits.cu), the lane emulator (tests/emu) runs it on
// the CPU against event streams recorded from the live reference.
#pragma once
#include "gb_common.cuh"

namespace gb {

constexpr int kSymPerBit = 20;          // constants.py:24 PSEUDOSYMBOLS_PER_NAVIGATION_BIT
constexpr int kPhaseWindow = 320;       // :135: the last 16 bits' worth of symbols vote on the phase
constexpr int kPhaseMinSymbols = 80;    // :75,:130: four bits' worth before a phase is chosen
constexpr int kResyncPeriod = 1000;     // config.py:38 x constants.py PSEUDOSYMBOLS_PER_SECOND
constexpr int kHealthBits = 10;         // config.py:41
constexpr int kQueueCap = 128;          // queued symbols (>= 81 + 20: first decision arrives with 81 queued)
constexpr double kResyncHorizon = 40.0; // :283: no phase changes after 40 s of receiver time

struct BitEvent {          // EmitNavigationBitEvent (:29-39), 32 bytes
    double receiver_timestamp;                // start of the bit's first pseudosymbol
    double trailing_edge_receiver_timestamp;  // end of its last one
    int ms_index;                             // symbol (within the call) whose arrival emitted the bit
    int bit_value;                            // 1 / 0 / -1 = BitValue.UNKNOWN
    int slide;                                // integrator.slide when the bit was emitted
    int pad_;
};
static_assert(sizeof(BitEvent) == 32, "bit event must stay 32 bytes");

// Scalar part of the integrator: lives in registers while the kernel walks a channel.
struct BitHead {
    long long processed;  // history.processed_pseudosymbol_count
    int mod_period;       // processed % kResyncPeriod and % kSymPerBit, kept incrementally
    int mod_bit;
    int prev_decision;    // history.previous_bit_phase_decision, -1 = None
    int determined;       // history.determined_bit_phase, -1 = None
    int slide;
    int cursor;           // history.pseudosymbol_cursor_within_queue (may go negative)
    int qlen, qhead;      // history.queued_pseudosymbols as a ring
    int seq_unknown;      // history.sequential_unknown_bit_value_counter
    int failed, emitted;  // history.failed_bit_count / emitted_bit_count
    int seen_count, seen_head;  // last kPhaseWindow symbols of history.last_seen_pseudosymbols
    int bits_count;       // emitted bits remembered (<= kHealthBits) ...
    int unknown_hist;     // ... bit k set: the k-th most recent one was BitValue.UNKNOWN (history.last_emitted_bits)
    int overflow;         // the queue overflowed: only reachable once no phase can ever be chosen again (:283)
    int stopped;          // the tracking channel lost lock: nothing further is integrated
    int pad_;
};

struct BitState {
    BitHead h;
    signed char seen[kPhaseWindow];
    signed char qsym[kQueueCap];
    double qstart[kQueueCap], qend[kQueueCap];
};

GB_HD inline void bit_state_init(BitState& st) {
    BitHead& h = st.h;
    h.processed = 0;
    h.mod_period = h.mod_bit = 0;
    h.prev_decision = h.determined = -1;
    h.slide = h.cursor = h.qlen = h.qhead = 0;
    h.seq_unknown = h.failed = h.emitted = 0;
    h.seen_count = h.seen_head = h.bits_count = h.unknown_hist = 0;
    h.overflow = h.stopped = h.pad_ = 0;
}

// :127-147 with _compute_bit_confidence_score (:105-125): the phase whose 20-symbol blocks agree most over the window;
// the first phase wins ties (max() over an insertion-ordered dict).  The score is the integer sum of |block sums|
// divided by constants, so the integers are compared.
GB_HD inline int bit_redetermine_phase(const BitHead& h, const signed char* seen) {
    if (h.seen_count < kPhaseMinSymbols) return -1;
    const int n = h.seen_count;  // <= kPhaseWindow
    const int first = (h.seen_head - n + 2 * kPhaseWindow) % kPhaseWindow;
    const int blocks = n / kSymPerBit;
    int best = 0, best_score = -1;
    for (int p = 0; p < kSymPerBit; ++p) {
        int score = 0;
        for (int b = 0; b < blocks; ++b) {
            int sum = 0;
            for (int j = 0; j < kSymPerBit; ++j) {
                int i = b * kSymPerBit + j + p;  // np.roll(values, -p)[b*20 + j]
                if (i >= n) i -= n;
                int pos = first + i;
                if (pos >= kPhaseWindow) pos -= kPhaseWindow;
                sum += seen[pos];
            }
            score += sum < 0 ? -sum : sum;
        }
        if (score > best_score) {
            best_score = score;
            best = p;
        }
    }
    return best;
}

// :217-246
GB_HD inline bool bit_should_resync(const BitHead& h) {
    if (h.mod_period == 0) return true;
    if (h.mod_bit != 0) return false;
    if (h.prev_decision < 0) return true;
    if (h.bits_count >= kHealthBits) {
        int unknown = 0;
        for (int k = 0; k < kHealthBits; ++k) unknown += (h.unknown_hist >> k) & 1;
        if ((static_cast<double>(unknown) / kHealthBits) * 100.0 >= 50.0) return true;  // config.py:43
    }
    return false;
}

// :248-276
GB_HD inline void bit_resync_if_necessary(BitHead& h, const signed char* seen) {
    if (!bit_should_resync(h)) return;
    const int before = h.prev_decision;
    const int after = bit_redetermine_phase(h, seen);
    h.prev_decision = after;
    h.determined = after;
    if (before < 0 && after >= 0) {
        if (after > 0) {
            h.cursor = after;
            h.slide = after;
        }
    } else if (before >= 0 && after >= 0 && before != after) {
        h.slide += after - before;
        h.cursor += after - before;
    }
}

// :163-192 on queue entries [at, at+20)
GB_HD inline void bit_emit(BitHead& h, const BitState& q, int at, int ms_index, BitEvent* out, int max_out, int& n_out) {
    int total = 0;
    for (int j = 0; j < kSymPerBit; ++j) total += q.qsym[(h.qhead + at + j) % kQueueCap];
    int value = total > 0 ? 1 : 0;
    const int conf = static_cast<int>((static_cast<double>(total) / kSymPerBit) * 100.0);  // :155-158
    if ((conf < 0 ? -conf : conf) <= 50) value = -1;
    h.unknown_hist = ((h.unknown_hist << 1) | (value < 0 ? 1 : 0)) & ((1 << kHealthBits) - 1);
    if (h.bits_count < kHealthBits) h.bits_count++;
    if (value < 0) {
        h.seq_unknown++;
        h.failed++;
        if (h.seq_unknown >= 30) h.determined = -1;  // :184-187 _reset_selected_bit_phase
    } else {
        h.seq_unknown = 0;
    }
    if (n_out < max_out) {
        BitEvent ev;
        ev.receiver_timestamp = q.qstart[(h.qhead + at) % kQueueCap];
        ev.trailing_edge_receiver_timestamp = q.qend[(h.qhead + at + kSymPerBit - 1) % kQueueCap];
        ev.ms_index = ms_index;
        ev.bit_value = value;
        ev.slide = h.slide;
        ev.pad_ = 0;
        out[n_out] = ev;
    }
    n_out++;  // counts past max_out so the caller can see the truncation
}

// process_pseudosymbol (:278-288) for one symbol of one channel.  `h` is the channel's scalar state (the caller may hold
// it in registers), `q` its arrays.
GB_HD inline void bit_step(BitHead& h, BitState& q, int symbol, double receiver_timestamp, double start, double end,
                           int ms_index, BitEvent* out, int max_out, int& n_out) {
    if (h.qlen == kQueueCap) {  // see `overflow`
        h.qhead = (h.qhead + 1) % kQueueCap;
        h.qlen--;
        h.overflow = 1;
    }
    const int slot = (h.qhead + h.qlen) % kQueueCap;
    q.qsym[slot] = static_cast<signed char>(symbol);
    q.qstart[slot] = start;
    q.qend[slot] = end;
    h.qlen++;
    q.seen[h.seen_head] = static_cast<signed char>(symbol);
    h.seen_head = h.seen_head + 1 == kPhaseWindow ? 0 : h.seen_head + 1;
    if (h.seen_count < kPhaseWindow) h.seen_count++;

    if (receiver_timestamp < kResyncHorizon) bit_resync_if_necessary(h, q.seen);

    if (h.determined >= 0) {  // :194-215
        // pending = queued[cursor:] with Python slice semantics
        const int from = h.cursor >= 0 ? (h.cursor < h.qlen ? h.cursor : h.qlen) : (h.qlen + h.cursor > 0 ? h.qlen + h.cursor : 0);
        const int whole = (h.qlen - from) / kSymPerBit;
        for (int c = 0; c < whole; ++c) {
            bit_emit(h, q, from + c * kSymPerBit, ms_index, out, max_out, n_out);
            h.cursor += kSymPerBit;
            h.emitted++;
        }
        if (h.qlen >= kSymPerBit) {
            const int unread = h.qlen - h.cursor;
            h.qhead = (h.qhead + h.qlen - kSymPerBit) % kQueueCap;
            h.qlen = kSymPerBit;
            h.cursor = kSymPerBit - unread;
        }
    }
    h.processed++;
    h.mod_period = h.mod_period + 1 == kResyncPeriod ? 0 : h.mod_period + 1;
    h.mod_bit = h.mod_bit + 1 == kSymPerBit ? 0 : h.mod_bit + 1;
}

}  // namespace gb
