// Tokeniser + padding collator: the host-side callers in front of zvx_encode.
//   Symbols                         zerovox/tts/symbols.py:2-49
//   ZeroVoxTTS.transcript2phonemids zerovox/tts/synthesize.py:145-190
//   collate_fn (pad_sequence + get_mask_from_lengths) zerovox/tts/data.py:56-60, 82-83; fs2.py:565-573
// Pure host C++ (no device work): the reference does this in Python per character; here it is one pass over the UTF-8
// bytes with vocabulary lookup tables, so a serving front-end can tokenise and collate a batch without the interpreter.
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/zerovox_b200.h"

namespace {

// Decodes one code point (Python iterates `str` by code point); malformed bytes decode to U+FFFD and advance by one.
inline uint32_t next_code_point(const unsigned char* s, size_t n, size_t& i) {
    const unsigned char c = s[i];
    auto cont = [&](size_t k) { return i + k < n && (s[i + k] & 0xC0) == 0x80; };
    if (c < 0x80) { i += 1; return c; }
    if ((c & 0xE0) == 0xC0 && cont(1)) { uint32_t v = ((c & 0x1F) << 6) | (s[i + 1] & 0x3F); i += 2; return v; }
    if ((c & 0xF0) == 0xE0 && cont(1) && cont(2)) {
        uint32_t v = ((c & 0x0F) << 12) | ((s[i + 1] & 0x3F) << 6) | (s[i + 2] & 0x3F); i += 3; return v;
    }
    if ((c & 0xF8) == 0xF0 && cont(1) && cont(2) && cont(3)) {
        uint32_t v = ((c & 0x07) << 18) | ((s[i + 1] & 0x3F) << 12) | ((s[i + 2] & 0x3F) << 6) | (s[i + 3] & 0x3F);
        i += 4; return v;
    }
    i += 1;
    return 0xFFFD;
}

struct Vocab {
    int32_t ascii[128];                              // -1 = absent
    std::unordered_map<uint32_t, int32_t> wide;
    int32_t count = 0;                               // distinct entries
    Vocab() { for (auto& a : ascii) a = -1; }
    void set(uint32_t cp, int32_t id) {              // later duplicates overwrite (dict semantics, symbols.py:11-13)
        if (find(cp) < 0) ++count;
        if (cp < 128) ascii[cp] = id; else wide[cp] = id;
    }
    int32_t find(uint32_t cp) const {
        if (cp < 128) return ascii[cp];
        auto it = wide.find(cp);
        return it == wide.end() ? -1 : it->second;
    }
};

}  // namespace

struct zvx_symbols {
    Vocab phones, puncts;
    std::string err;
};

static std::string g_sym_error;

extern "C" {

int zvx_symbols_create(const char* phones_utf8, const char* puncts_utf8, zvx_symbols** out) {
    if (!phones_utf8 || !puncts_utf8 || !out) { g_sym_error = "zvx_symbols_create: null argument"; return -1; }
    try {
        std::unique_ptr<zvx_symbols> s(new zvx_symbols());
        const auto* p = reinterpret_cast<const unsigned char*>(phones_utf8);
        size_t n = std::strlen(phones_utf8), i = 0;
        int32_t id = 0;
        while (i < n) s->phones.set(next_code_point(p, n, i), id++);
        p = reinterpret_cast<const unsigned char*>(puncts_utf8);
        n = std::strlen(puncts_utf8); i = 0; id = 1;                 // id 0 is the reserved '_NP_' (symbols.py:15-16)
        while (i < n) s->puncts.set(next_code_point(p, n, i), id++);
        *out = s.release();
        return 0;
    } catch (const std::exception& e) {
        g_sym_error = e.what();
        return 1;
    }
}

void zvx_symbols_destroy(zvx_symbols* s) { delete s; }

const char* zvx_symbols_last_error(const zvx_symbols* s) { return s ? s->err.c_str() : g_sym_error.c_str(); }

int zvx_symbols_num_phones(const zvx_symbols* s) { return s ? s->phones.count : -1; }
int zvx_symbols_num_puncts(const zvx_symbols* s) { return s ? s->puncts.count + 1 : -1; }   // + '_NP_'

int zvx_transcript2phonemids(zvx_symbols* s, const char* transcript_utf8, int32_t* phone_ids, int32_t* punct_ids,
                             int capacity) {
    if (!s) return -1;
    if (!transcript_utf8 || capacity < 0 || (capacity > 0 && (!phone_ids || !punct_ids))) {
        s->err = "zvx_transcript2phonemids: null argument";
        return -1;
    }
    const auto* t = reinterpret_cast<const unsigned char*>(transcript_utf8);
    const size_t n = std::strlen(transcript_utf8);
    size_t i = 0;
    int count = 0;
    int32_t punct = 0;
    int32_t last_punct = 0;
    bool in_run = false;
    while (i < n) {
        const uint32_t cp = next_code_point(t, n, i);
        const int32_t pu = s->puncts.find(cp);
        if (cp == ' ' || pu >= 0) {                                  // synthesize.py:157: blank or punctuation
            if (pu < 0) {                                            // encode_punct(' ') raises KeyError there
                s->err = "zvx_transcript2phonemids: ' ' is not in the punctuation vocabulary (reference raises KeyError)";
                return -2;
            }
            if (pu > punct) punct = pu;
            in_run = true;
            last_punct = punct;
            continue;
        }
        if (in_run) {                                                // the run ended: it labels the preceding phone
            if (count > 0 && count <= capacity) punct_ids[count - 1] = last_punct;
            in_run = false;
        }
        const int32_t ph = s->phones.find(cp);
        if (ph < 0) continue;                                        // neither: skipped (synthesize.py:181-183)
        punct = 0;
        if (count < capacity) { phone_ids[count] = ph; punct_ids[count] = 0; }
        ++count;
    }
    if (in_run && count > 0 && count <= capacity) punct_ids[count - 1] = last_punct;
    if (count > capacity) {
        s->err = "zvx_transcript2phonemids: capacity too small; the return value is the needed size";
    }
    return count;                                                    // > capacity: nothing beyond capacity was written
}

int zvx_collate(const int32_t* const* phone_seqs, const int32_t* const* punct_seqs, const int32_t* lens, int B, int T,
                int32_t* phoneme, int32_t* puncts, uint8_t* phoneme_mask) {
    if (B < 0 || T < 0 || (B > 0 && (!phone_seqs || !punct_seqs || !lens))) return -1;
    if (B > 0 && T > 0 && (!phoneme || !puncts)) return -1;
    for (int b = 0; b < B; ++b) {
        const int len = lens[b];
        if (len < 0 || len > T || (len > 0 && (!phone_seqs[b] || !punct_seqs[b]))) return -2;
        int32_t* ph = phoneme + (size_t)b * T;
        int32_t* pu = puncts + (size_t)b * T;
        std::memcpy(ph, phone_seqs[b], (size_t)len * sizeof(int32_t));
        std::memcpy(pu, punct_seqs[b], (size_t)len * sizeof(int32_t));
        std::memset(ph + len, 0, (size_t)(T - len) * sizeof(int32_t));         // pad_sequence pads with 0
        std::memset(pu + len, 0, (size_t)(T - len) * sizeof(int32_t));
        if (phoneme_mask) {                                                     // get_mask_from_lengths: ids >= len
            std::memset(phoneme_mask + (size_t)b * T, 0, (size_t)len);
            std::memset(phoneme_mask + (size_t)b * T + len, 1, (size_t)(T - len));
        }
    }
    return 0;
}

}  // extern "C"
