// vgs_tree.cu -- native `.tree` ingest: the text of a classifier / PST `.tree` file straight to the flat
// context table that vgs_load_models takes, without the JSON round trip.  Host code only.
//
// Reference behaviour restated (src/parse_trees_to_json.py:18-55):
//   :23      only lines whose first five characters are "Node:" count
//   :36      numbers = re.findall('-?[0-9]+\.?[0-9]*', line)       -- maximal tokens: optional '-', digits, optional '.', digits
//   :39      total = int(numbers[6]) + ... + int(numbers[9])        -- the four "next symbol" counts (A, C, G, T order)
//   :40-44   deltas: the row is float(numbers[14..17])
//   :45-49   else:   the row is int(numbers[k]) / total             -- Python true division of two ints
//   :51-54   key = second match of '[A-Za-z#]+' (the first is "Node"); '#' is the root context ''
//   :24-29   tree[key] = row (a repeated key overwrites the row, keeps its first position: dict semantics);
//            occurrence_probability[key] = total / (total of the FIRST node line - max(len(key) - 1, 0))
#include <stdlib.h>
#include <string.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "vgs_common.cuh"

namespace {

struct Token {
    const char *p;
    int len;
    bool has_dot;
};

inline bool is_digit(char c) { return c >= '0' && c <= '9'; }
inline bool is_word(char c) { return (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '#'; }

// re.findall('-?[0-9]+\.?[0-9]*', line): leftmost, non-overlapping, greedy
void find_numbers(const char *b, const char *e, std::vector<Token> *out) {
    out->clear();
    const char *p = b;
    while (p < e) {
        const char *q = p;
        if (*q == '-' && q + 1 < e && is_digit(q[1])) ++q;
        if (!is_digit(*q)) { ++p; continue; }
        while (q < e && is_digit(*q)) ++q;
        bool dot = false;
        if (q < e && *q == '.') { dot = true; ++q; }
        while (q < e && is_digit(*q)) ++q;
        out->push_back(Token{p, (int)(q - p), dot});
        p = q;
    }
}

// Python int(token): digits with an optional sign; a '.' in the token is a ValueError there
bool to_int(const Token &t, long long *v) {
    if (t.has_dot || t.len > 18) return false;
    char buf[24];
    memcpy(buf, t.p, (size_t)t.len);
    buf[t.len] = 0;
    *v = strtoll(buf, nullptr, 10);
    return true;
}

// Python float(token): correctly rounded decimal -> binary64, as strtod
double to_float(const Token &t) {
    char buf[64];
    const int n = t.len < 63 ? t.len : 63;
    memcpy(buf, t.p, (size_t)n);
    buf[n] = 0;
    return strtod(buf, nullptr);
}

}  // namespace

extern "C" int vgs_parse_tree_text(const char *text, int64_t n_bytes, int deltas, int64_t capacity, uint8_t *ctx_len_out,
                                   uint32_t *ctx_code_out, double *prob_out, double *occurrence_out, int64_t *n_ctx_out) {
    VGS_CHECK_ARG(text && n_bytes >= 0 && n_ctx_out, "vgs_parse_tree_text: bad arguments");
    std::unordered_map<std::string, int64_t> index;
    std::vector<Token> numbers;
    int64_t n = 0;
    long long root_total = -1;
    bool first = true;
    const char *p = text, *end = text + n_bytes;
    while (p < end) {
        const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *le = eol ? eol : end;
        if (le - p >= 5 && memcmp(p, "Node:", 5) == 0) {
            find_numbers(p, le, &numbers);
            if (numbers.size() < (deltas ? 18u : 10u)) {
                vgs_set_error("vgs_parse_tree_text: a Node line has %d numbers, %d are needed (IndexError in the reference)",
                              (int)numbers.size(), deltas ? 18 : 10);
                return VGS_ERR_INVALID;
            }
            long long c[4], total = 0;
            for (int k = 0; k < 4; ++k) {
                if (!to_int(numbers[(size_t)(6 + k)], &c[k])) {
                    vgs_set_error("vgs_parse_tree_text: count is not an integer (ValueError in the reference)");
                    return VGS_ERR_INVALID;
                }
                total += c[k];
            }
            double row[4];
            if (deltas) {
                for (int k = 0; k < 4; ++k) row[k] = to_float(numbers[(size_t)(14 + k)]);
            } else {
                if (total == 0) {
                    vgs_set_error("vgs_parse_tree_text: a node has no next-symbol counts (ZeroDivisionError in the reference)");
                    return VGS_ERR_INVALID;
                }
                for (int k = 0; k < 4; ++k) row[k] = (double)c[k] / (double)total;
            }
            // key: the second run of [A-Za-z#]
            const char *q = p;
            const char *kb = nullptr, *ke = nullptr;
            for (int run = 0; run < 2 && q < le; ++run) {
                while (q < le && !is_word(*q)) ++q;
                kb = q;
                while (q < le && is_word(*q)) ++q;
                ke = q;
                if (kb == ke) { kb = ke = nullptr; break; }
            }
            if (!kb) {
                vgs_set_error("vgs_parse_tree_text: a Node line has no context field (IndexError in the reference)");
                return VGS_ERR_INVALID;
            }
            std::string key(kb, (size_t)(ke - kb));
            if (key == "#") key.clear();
            if (first) { root_total = total; first = false; }
            const long long denom = root_total - (long long)(key.size() > 1 ? key.size() - 1 : 0);
            if (denom == 0) {
                vgs_set_error("vgs_parse_tree_text: occurrence denominator is zero (ZeroDivisionError in the reference)");
                return VGS_ERR_INVALID;
            }
            const double occ = (double)total / (double)denom;
            auto it = index.find(key);
            int64_t at;
            if (it != index.end()) {
                at = it->second;
            } else {
                at = n++;
                index.emplace(key, at);
                if (at < capacity) {
                    if (key.size() > (size_t)VGS_MAX_ORDER) {
                        vgs_set_error("context longer than %d symbols is not supported", VGS_MAX_ORDER);
                        return VGS_ERR_UNSUPPORTED;
                    }
                    uint32_t code = 0;
                    for (char ch : key) {
                        uint32_t x;
                        if (ch == 'A') x = 0; else if (ch == 'C') x = 1; else if (ch == 'G') x = 2; else if (ch == 'T') x = 3;
                        else {
                            vgs_set_error("'%c'", ch);   // KeyError(symbol) where the reference indexes the alphabet (src/vlmc/vlmc.pyx:125)
                            return VGS_ERR_SYMBOL;
                        }
                        code = (code << 2) | x;
                    }
                    if (ctx_len_out) ctx_len_out[at] = (uint8_t)key.size();
                    if (ctx_code_out) ctx_code_out[at] = code;
                }
            }
            if (at < capacity) {
                if (prob_out) memcpy(prob_out + 4 * at, row, sizeof(row));
                if (occurrence_out) occurrence_out[at] = occ;
            }
        }
        p = eol ? eol + 1 : end;
    }
    *n_ctx_out = n;
    if (n > capacity && (ctx_len_out || ctx_code_out || prob_out || occurrence_out)) {
        vgs_set_error("vgs_parse_tree_text: %lld contexts, capacity %lld", (long long)n, (long long)capacity);
        return VGS_ERR_MEMORY;
    }
    return VGS_OK;
}
