/*
 * rure.h -- C ABI of the regex_b200 search backend.
 *
 * Drop-in boundary: these are the entry points a C caller (or the reference's
 * own FFI crate) binds for the search path.  Names, argument meaning, ownership
 * and error behaviour follow the reference's C API, regex-capi/include/rure.h;
 * each declaration cites the reference line it replaces.  All matching runs on
 * the GPU (CUDA, sm_100a); there is no CPU matching path behind this header.
 *
 * Differences from the reference, all explicit:
 *   - rure_compile fails (NULL + message) for patterns that need Unicode-aware
 *     word boundaries (\b, \B without (?-u)) and for patterns whose dense DFA
 *     exceeds the table budget ("exceeds size limit").
 *   - capture groups other than group 0 are out of scope: rure_captures_len
 *     reports 1 and only index 0 is ever populated.
 */
#ifndef REGEX_B200_RURE_H
#define REGEX_B200_RURE_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rure rure;                 /* reference rure.h:29  (bytes::Regex)    */
typedef struct rure_set rure_set;         /* reference rure.h:36  (bytes::RegexSet) */
typedef struct rure_options rure_options; /* reference rure.h:47 */

/* reference rure.h:56-68 */
#define RURE_FLAG_CASEI (1 << 0)
#define RURE_FLAG_MULTI (1 << 1)
#define RURE_FLAG_DOTNL (1 << 2)
#define RURE_FLAG_SWAP_GREED (1 << 3)
#define RURE_FLAG_SPACE (1 << 4)
#define RURE_FLAG_UNICODE (1 << 5)
#define RURE_DEFAULT_FLAGS RURE_FLAG_UNICODE

/* reference rure.h:73-78: byte offsets, start inclusive, end exclusive */
typedef struct rure_match {
  size_t start;
  size_t end;
} rure_match;

typedef struct rure_captures rure_captures;                     /* reference rure.h:95  */
typedef struct rure_iter rure_iter;                             /* reference rure.h:106 */
typedef struct rure_iter_capture_names rure_iter_capture_names; /* reference rure.h:117 */
typedef struct rure_error rure_error;                           /* reference rure.h:130 */

/* reference rure.h:147 -- aborts the process on error, like the reference */
rure *rure_compile_must(const char *pattern);
/* reference rure.h:168-170 -- pattern is UTF-8, not NUL terminated */
rure *rure_compile(const uint8_t *pattern, size_t length, uint32_t flags, rure_options *options,
                   rure_error *error);
/* reference rure.h:177 */
void rure_free(rure *re);
/* reference rure.h:197-198 -- Regex::is_match_at (src/re_bytes.rs:587-589) */
bool rure_is_match(rure *re, const uint8_t *haystack, size_t length, size_t start);
/* reference rure.h:220-221 -- Regex::find_at (src/re_bytes.rs:603-605) */
bool rure_find(rure *re, const uint8_t *haystack, size_t length, size_t start, rure_match *match);
/* reference rure.h:248-249 -- group 0 only */
bool rure_find_captures(rure *re, const uint8_t *haystack, size_t length, size_t start,
                        rure_captures *captures);
/* reference rure.h:273-274 -- Regex::shortest_match_at (src/re_bytes.rs:572-577) */
bool rure_shortest_match(rure *re, const uint8_t *haystack, size_t length, size_t start, size_t *end);
/* reference rure.h:285 -- always -1 (named groups are out of scope) */
int32_t rure_capture_name_index(rure *re, const char *name);
/* reference rure.h:292-307 -- yields nothing */
rure_iter_capture_names *rure_iter_capture_names_new(rure *re);
void rure_iter_capture_names_free(rure_iter_capture_names *it);
bool rure_iter_capture_names_next(rure_iter_capture_names *it, char **name);
/* reference rure.h:317-345 -- find_iter (src/re_trait.rs:197-220, regex-capi/src/rure.rs:323-361).
 * The haystack passed to successive rure_iter_next calls must stay unchanged;
 * all matches are computed on the GPU on the first call and then replayed. */
rure_iter *rure_iter_new(rure *re);
void rure_iter_free(rure_iter *it);
bool rure_iter_next(rure_iter *it, const uint8_t *haystack, size_t length, rure_match *match);
/* reference rure.h:369-371 -- group 0 only */
bool rure_iter_next_captures(rure_iter *it, const uint8_t *haystack, size_t length,
                             rure_captures *captures);
/* reference rure.h:385-411 */
rure_captures *rure_captures_new(rure *re);
void rure_captures_free(rure_captures *captures);
bool rure_captures_at(rure_captures *captures, size_t i, rure_match *match);
size_t rure_captures_len(rure_captures *captures);
/* reference rure.h:423-454 */
rure_options *rure_options_new(void);
void rure_options_free(rure_options *options);
void rure_options_size_limit(rure_options *options, size_t limit);
/* Reinterpreted: the dense transition-table budget is 16x this value. */
void rure_options_dfa_size_limit(rure_options *options, size_t limit);
/* reference rure.h:476-481 */
rure_set *rure_compile_set(const uint8_t **patterns, const size_t *patterns_lengths,
                           size_t patterns_count, uint32_t flags, rure_options *options,
                           rure_error *error);
/* reference rure.h:488 */
void rure_set_free(rure_set *re);
/* reference rure.h:505-506 -- RegexSet::is_match_at (src/re_set.rs:145) */
bool rure_set_is_match(rure_set *re, const uint8_t *haystack, size_t length, size_t start);
/* reference rure.h:532-533 -- RegexSet::read_matches_at (src/re_set.rs:206-213);
 * matches[0..rure_set_len) is zeroed first (regex-capi/src/rure.rs:560-563) */
bool rure_set_matches(rure_set *re, const uint8_t *haystack, size_t length, size_t start,
                      bool *matches);
/* reference rure.h:538 */
size_t rure_set_len(rure_set *re);
/* reference rure.h:551-568 */
rure_error *rure_error_new(void);
void rure_error_free(rure_error *err);
const char *rure_error_message(rure_error *err);

#ifdef __cplusplus
}
#endif
#endif
