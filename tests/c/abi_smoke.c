/* The C ABI used from plain C (no Python, no torch): version string, the host-side formatter on a tiny
 * batch, and the error convention (a negative EMDR2_E* code plus a message) for a call that cannot
 * succeed without a device or with bad arguments.  Built and run by tests/test_abi.py. */
#include <stdio.h>
#include <string.h>

#include "emdr2_b200.h"

int main(void) {
  const char* v = emdr2_version();
  if (!v || strncmp(v, "emdr2_b200", 10) != 0) return 10;

  /* one question (uid -1, tokens 7 8), two candidates; the second has two passages and is the first of them */
  int64_t uid[1] = {-1}, query[4] = {7, 8, 0, 0}, qlen[1] = {2};
  int32_t cand_begin[2] = {0, 2};
  int64_t cand_id[2] = {11, 12};
  int32_t meta[12] = {1, 1, 0, 3, 0, 0, /* title_len n_docs main_idx doc_len[3] */
                      2, 2, 0, 2, 2, 0};
  int64_t tokens[] = {40, 41, 42, 43, /* cand 0: title 40 | doc 41 42 43 */
                      50, 51, 60, 61, 70, 71 /* cand 1: title 50 51 | doc0 60 61 | doc1 70 71 */};
  int64_t ctx[2 * 8], typ[2 * 8], ext[2 * 12], one[2 * 12];
  int32_t longest[3], row_len[6];
  int rc = emdr2_format_passages(1, 2, uid, query, 4, qlen, cand_begin, cand_id, meta, tokens, 10, 8, 12, 101, 102,
                                 0, ctx, typ, ext, one, longest, row_len);
  if (rc != EMDR2_OK) {
    fprintf(stderr, "format failed: %s\n", emdr2_last_error());
    return 11;
  }
  const int64_t want_ctx0[8] = {101, 40, 102, 41, 42, 43, 102, 0};
  const int64_t want_ext1[12] = {7, 8, 50, 51, 102, 60, 61, 70, 71, 102, 0, 0};
  const int64_t want_one1[12] = {7, 8, 50, 51, 102, 60, 61, 102, 0, 0, 0, 0};
  if (memcmp(ctx, want_ctx0, sizeof want_ctx0) || memcmp(ext + 12, want_ext1, sizeof want_ext1) ||
      memcmp(one + 12, want_one1, sizeof want_one1))
    return 12;
  if (longest[0] != 7 || longest[1] != 10 || longest[2] != 8 || row_len[3] != 10) return 13;

  /* error convention: bad sizes -> EMDR2_EINVAL and a message; nothing aborts */
  rc = emdr2_format_passages(1, 2, uid, query, 4, qlen, cand_begin, cand_id, meta, tokens, 10, 1, 12, 101, 102, 0,
                             ctx, typ, ext, one, longest, row_len);
  if (rc != EMDR2_EINVAL || strlen(emdr2_last_error()) == 0) return 14;
  void* handle = NULL;
  rc = emdr2_mips_create(7, EMDR2_DTYPE_FP16, 0, &handle); /* d must be a multiple of 8 */
  if (rc == EMDR2_OK || handle != NULL) return 15;
  printf("%s\n", v);
  return 0;
}
