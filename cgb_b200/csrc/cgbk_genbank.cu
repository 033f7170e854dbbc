// GenBank feature-table index (host only): the structural scan of "FEATURES ... ORIGIN" that
// cgb_b200/genbank.py used to do line by line in Python -- the remaining host bottleneck of the
// per-genome loop once the sequence goes through bytes.translate (the reference
// leaves all of it to Biopython's SeqIO.read, cgb/chromid.py:32-35).
//
// One pass over the block yields, per feature: the key span, the location span and the parsed
// location (start, end, strand, compound -- Biopython's FeatureLocation as Chromid.genes reads it,
// cgb/chromid.py:107-117), and the range of its qualifiers; per qualifier: the key span and the RAW
// value span (continuation lines and quotes still in it).  Python decodes only the handful of values
// the gene table needs.  Layout rules (INSDC feature table): feature key in column 6, location and
// qualifiers from column 22, a qualifier starts with '/', a quoted value runs until a quote that
// closes it (an even number of quotes, the last one at the end of a line).
#include <stdint.h>
#include <string.h>

#include "cgbk_internal.cuh"

namespace {

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\f' || c == '\v'; }

// "^ {5}\S": a feature key line
inline bool feature_start(const char* p, int64_t len) {
    return len > 5 && p[0] == ' ' && p[1] == ' ' && p[2] == ' ' && p[3] == ' ' && p[4] == ' ' && !is_space(p[5]);
}

// "^ {21}/": a qualifier line
inline bool qualifier_start(const char* p, int64_t len) {
    if (len < 22) return false;
    for (int i = 0; i < 21; ++i)
        if (p[i] != ' ') return false;
    return p[21] == '/';
}

// ---- location: genbank.parse_location restated over a whitespace-free copy -------------------
struct Loc { int32_t start, end, strand, compound; };

bool starts_with(const char* s, int64_t n, const char* w) {
    const int64_t k = (int64_t)strlen(w);
    return n >= k && memcmp(s, w, (size_t)k) == 0;
}

// digits at s[i..): value (saturating at int32 max) and the index after them; false if none
bool read_int(const char* s, int64_t n, int64_t& i, int64_t& v) {
    const int64_t i0 = i;
    v = 0;
    while (i < n && s[i] >= '0' && s[i] <= '9') {
        if (v < (1ll << 40)) v = v * 10 + (s[i] - '0');
        ++i;
    }
    return i > i0;
}

inline bool ref_char(char c) {
    return (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || (c >= '0' && c <= '9') || c == '_' || c == '.';
}

Loc parse_location(const char* raw, int64_t len, char* tmp) {
    int64_t n = 0;
    for (int64_t i = 0; i < len; ++i)
        if (!is_space(raw[i])) tmp[n++] = raw[i];
    const char* s = tmp;
    Loc L = {0, 0, 1, 0};
    while (starts_with(s, n, "complement(") && n > 0 && s[n - 1] == ')') {
        s += 11; n -= 12;
        L.strand = -L.strand;
    }
    if (starts_with(s, n, "join(") || starts_with(s, n, "order(") || starts_with(s, n, "bond(")) {
        // hull of every number that is not part of an "accession:" prefix
        int64_t lo = -1, hi = -1, i = 0;
        while (i < n) {
            if (ref_char(s[i])) {
                // a run of reference characters: a number, or an accession when followed by ':'
                int64_t j = i;
                while (j < n && ref_char(s[j])) ++j;
                if (j < n && s[j] == ':') { i = j + 1; continue; }
                // numbers inside the run (digits separated by '.' etc.)
                int64_t k = i;
                while (k < j) {
                    if (s[k] >= '0' && s[k] <= '9') {
                        int64_t v;
                        read_int(s, j, k, v);
                        if (lo < 0 || v < lo) lo = v;
                        if (v > hi) hi = v;
                    } else {
                        ++k;
                    }
                }
                i = j;
            } else {
                ++i;
            }
        }
        L.start = lo < 0 ? 0 : (int32_t)(lo - 1);
        L.end = hi < 0 ? 0 : (int32_t)hi;
        L.compound = 1;
        return L;
    }
    // ^(?:[A-Za-z0-9_.]+:)?[<>]?(\d+)(?:\.\.[<>]?(\d+)|\.[<>]?(\d+)|\^(\d+))?$
    int64_t i = 0;
    {
        int64_t j = 0;
        while (j < n && ref_char(s[j])) ++j;
        if (j > 0 && j < n && s[j] == ':') i = j + 1;
    }
    if (i < n && (s[i] == '<' || s[i] == '>')) ++i;
    int64_t a, b;
    if (!read_int(s, n, i, a)) { L.compound = 1; return L; }
    if (i == n) { L.start = (int32_t)(a - 1); L.end = (int32_t)a; return L; }              // single base
    if (i + 1 < n && s[i] == '.' && s[i + 1] == '.') {
        i += 2;
        if (i < n && (s[i] == '<' || s[i] == '>')) ++i;
        if (read_int(s, n, i, b) && i == n) { L.start = (int32_t)(a - 1); L.end = (int32_t)b; return L; }
    } else if (s[i] == '.') {
        ++i;
        if (i < n && (s[i] == '<' || s[i] == '>')) ++i;
        if (read_int(s, n, i, b) && i == n) { L.start = (int32_t)(a - 1); L.end = (int32_t)b; return L; }
    } else if (s[i] == '^') {
        ++i;
        if (read_int(s, n, i, b) && i == n) { L.start = (int32_t)a; L.end = (int32_t)a; return L; }
    }
    L.start = 0; L.end = 0; L.compound = 1;                                                  // no match
    return L;
}

}  // namespace

// feat[k][10] = key_off, key_len, loc_off, loc_len, q_first, q_count, start, end, strand, compound
// qual[j][5]  = key_off, key_len, val_off, val_len, simple   (val_len = -1: a flag qualifier without
//               '='; simple = 1: the raw value is one line, quoted, with no quote inside)
// Offsets are relative to `block`.  Returns CGBK_OK and the counts; when a capacity is too small the
// counts still say what is needed (nothing beyond the capacity is written).
extern "C" int cgbk_genbank_index(const char* block, int64_t n, int32_t* feat, int32_t feat_cap, int32_t* qual,
                                  int32_t qual_cap, int32_t* n_feat, int32_t* n_qual) {
    if (!block || n < 0 || !n_feat || !n_qual || (feat_cap > 0 && !feat) || (qual_cap > 0 && !qual))
        return cgbk_set_error(CGBK_ERR_INVALID, "NULL or empty argument");
    if (n >= (1ll << 31)) return cgbk_set_error(CGBK_ERR_UNSUPPORTED, "feature table of 2 GB or more");
    int32_t nf = 0, nq = 0;
    // state of the item being extended by continuation lines
    enum { NONE, LOCATION, VALUE } open = NONE;
    int64_t item_off = 0, item_end = 0;      // span of the open location / value (end = last line's end)
    bool quoted = false, multi_line = false;
    int64_t quotes = 0;                      // quotes in the open value so far
    char* tmp = (char*)malloc((size_t)(n > 0 ? n : 1));
    if (!tmp) return cgbk_set_error(CGBK_ERR_INVALID, "out of memory");

    auto rstrip_end = [&](int64_t off, int64_t end) {
        while (end > off && is_space(block[end - 1])) --end;
        return end;
    };
    auto close_item = [&]() {
        if (open == LOCATION && nf >= 1 && nf <= feat_cap) {
            int32_t* f = feat + 10 * (int64_t)(nf - 1);
            // the head as Python sees it: everything after the key up to the first qualifier,
            // right-stripped with the feature chunk when no qualifier follows
            f[2] = (int32_t)item_off; f[3] = (int32_t)(item_end - item_off);
            const Loc L = parse_location(block + item_off, item_end - item_off, tmp);
            f[6] = L.start; f[7] = L.end; f[8] = L.strand; f[9] = L.compound;
        } else if (open == VALUE && nq >= 1 && nq <= qual_cap) {
            int32_t* q = qual + 5 * (int64_t)(nq - 1);
            const int64_t e = rstrip_end(item_off, item_end);
            q[2] = (int32_t)item_off; q[3] = (int32_t)(e - item_off);          // right-stripped
            q[4] = (quoted && quotes == 2 && e - item_off >= 2 && block[e - 1] == '"' && !multi_line) ? 1 : 0;
        }
        open = NONE;
    };
    // a quoted value is complete once, right-stripped, it ends with a quote and holds an even number
    auto quote_closed = [&]() {
        const int64_t e = rstrip_end(item_off, item_end);
        return e - item_off >= 2 && block[e - 1] == '"' && (quotes % 2) == 0;
    };

    int64_t pos = 0;
    while (pos < n) {
        const char* nl = (const char*)memchr(block + pos, '\n', (size_t)(n - pos));
        const int64_t line_end = nl ? (int64_t)(nl - block) : n;        // exclusive, without the newline
        const char* p = block + pos;
        const int64_t len = line_end - pos;
        // anything after the table that starts in column 1 (CONTIG, BASE COUNT, WGS, COMMENT) is not a
        // feature: the table ends there (the same four words genbank.parse cuts at)
        if (len > 0 && p[0] != ' ' &&
            (starts_with(p, len, "CONTIG") || starts_with(p, len, "BASE COUNT") || starts_with(p, len, "WGS") ||
             starts_with(p, len, "COMMENT")))
            break;
        if (feature_start(p, len)) {
            close_item();
            ++nf;
            int64_t k = 5;
            while (k < len && p[k] != ' ') ++k;
            if (nf <= feat_cap) {
                int32_t* f = feat + 10 * (int64_t)(nf - 1);
                f[0] = (int32_t)(pos + 5); f[1] = (int32_t)(k - 5);
                f[2] = (int32_t)(pos + k); f[3] = 0; f[4] = nq; f[5] = 0;
                f[6] = 0; f[7] = 0; f[8] = 1; f[9] = 1;
            }
            open = LOCATION;
            item_off = pos + k; item_end = line_end;
        } else if (nf > 0 && qualifier_start(p, len) && !(open == VALUE && quoted && !quote_closed())) {
            close_item();
            ++nq;
            if (nf <= feat_cap) feat[10 * (int64_t)(nf - 1) + 5] += 1;
            int64_t k = 22;
            while (k < len && p[k] != '=') ++k;
            if (nq <= qual_cap) {
                int32_t* q = qual + 5 * (int64_t)(nq - 1);
                q[0] = (int32_t)(pos + 22);
                if (k < len) { q[1] = (int32_t)(k - 22); q[2] = (int32_t)(pos + k + 1); q[3] = 0; q[4] = 0; }
                else { q[1] = (int32_t)(len - 22); q[2] = (int32_t)line_end; q[3] = -1; q[4] = 0; }
            }
            if (k < len) {
                open = VALUE;
                item_off = pos + k + 1; item_end = line_end;
                quoted = item_off < line_end && block[item_off] == '"';
                multi_line = false;
                quotes = 0;
                for (int64_t i = item_off; i < line_end; ++i) quotes += block[i] == '"';
            } else {
                open = NONE;        // flag qualifier; the key may still be continued, which INSDC does not do
            }
        } else if (open != NONE) {
            // continuation line of the open location / value (also a "/..." line inside open quotes)
            if (open == VALUE) {
                for (int64_t i = pos; i < line_end; ++i) quotes += block[i] == '"';
                multi_line = true;
            }
            item_end = line_end;
        }
        pos = line_end + 1;
    }
    close_item();
    free(tmp);
    *n_feat = nf; *n_qual = nq;
    return CGBK_OK;
}

// ORIGIN block -> sequence bytes (host only).  `text + from` is the first line AFTER the "ORIGIN"
// line; lines of position number + blocks of bases run until a line starting with "//" or the end
// of the text.  Every byte that is not one of " 0123456789\t\r\n" is kept, ASCII lower case folded
// to upper -- what genbank.parse did with bytes.translate, in one pass that also finds the end of
// the record.  Standard lines ("%9d" + 6 x (" " + 10 letters) + newline, 76 bytes) take a fixed-layout
// path.  Returns the number of bases written to `out` (capacity `cap`; CGBK_ERR_INVALID when it does
// not fit) and, in *end_pos, the offset of the "//" line (or n).
extern "C" int cgbk_genbank_origin(const char* text, int64_t n, int64_t from, uint8_t* out, int64_t cap,
                                   int64_t* n_bases, int64_t* end_pos) {
    if (!text || !out || !n_bases || !end_pos || from < 0 || from > n)
        return cgbk_set_error(CGBK_ERR_INVALID, "NULL or out-of-range argument");
    static unsigned char lut[256];
    static bool ready = false;
    if (!ready) {
        unsigned char t[256];
        for (int c = 0; c < 256; ++c) t[c] = (unsigned char)((c >= 'a' && c <= 'z') ? c - 32 : c);
        const char* drop = " 0123456789\t\r\n";
        for (const char* d = drop; *d; ++d) t[(unsigned char)*d] = 0;
        t[0] = 0;                       // (a NUL byte cannot be told from "dropped"; GenBank text has none)
        memcpy(lut, t, 256);
        ready = true;                   // benign race: every thread writes the same table
    }
    const unsigned char* p = (const unsigned char*)text;
    int64_t pos = from, k = 0;
    while (pos < n) {
        if (p[pos] == '/' && pos + 1 < n && p[pos + 1] == '/') break;
        // fixed-layout line?
        if (pos + 76 <= n && p[pos + 75] == '\n' && p[pos + 9] == ' ' && p[pos + 20] == ' ' && p[pos + 31] == ' ' &&
            p[pos + 42] == ' ' && p[pos + 53] == ' ' && p[pos + 64] == ' ' && k + 60 <= cap) {
            unsigned char all = 0xff;
            unsigned char* o = out + k;
            bool lead_ok = true;
            for (int i = 0; i < 9; ++i) lead_ok &= lut[p[pos + i]] == 0;
            if (lead_ok) {
                for (int b = 0; b < 6; ++b) {
                    const unsigned char* q = p + pos + 10 + 11 * b;
                    for (int i = 0; i < 10; ++i) {
                        const unsigned char c = lut[q[i]];
                        o[10 * b + i] = c;
                        all &= (unsigned char)(c != 0 ? 0xff : 0);
                    }
                }
                if (all) { k += 60; pos += 76; continue; }
            }
        }
        // general line: byte by byte up to the newline
        while (pos < n && p[pos] != '\n') {
            const unsigned char c = lut[p[pos]];
            if (c) {
                if (k >= cap) return cgbk_set_error(CGBK_ERR_INVALID, "sequence buffer of %lld bytes is too small", (long long)cap);
                out[k++] = c;
            }
            ++pos;
        }
        if (pos < n) ++pos;             // the newline
    }
    *n_bases = k;
    *end_pos = pos;
    return CGBK_OK;
}
