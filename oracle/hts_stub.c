/* TEST INFRASTRUCTURE — not product code.
 *
 * Link-time stand-in for the eight htslib entry points the reference's
 * Mapping.cpp references (src/Mapping.cpp:44, :655-662, :739, :755, :810).
 * They are only reached with `-bo` (BAM output), which is outside the hot-path
 * scope (htslib is "OUT OF SCOPE — third-party writer").  The SAM
 * text path (`-o`) never calls them, so the reference binaries built by
 * oracle/Makefile behave exactly like a full build for every run we make;
 * asking them for BAM aborts loudly instead of silently writing nothing.
 */
#include <stdio.h>
#include <stdlib.h>

static void *die(const char *fn)
{
    fprintf(stderr, "oracle/_ref: %s() called — BAM output (-bo) is not built into the "
                    "oracle binaries; use -o (SAM)\n", fn);
    abort();
    return NULL;
}

void *bam_init1(void) { return die("bam_init1"); }
void bam_destroy1(void *b) { (void)b; die("bam_destroy1"); }
void *hts_open_format(const char *fn, const char *mode, const void *fmt)
{ (void)fn; (void)mode; (void)fmt; return die("hts_open_format"); }
int hts_close(void *fp) { (void)fp; die("hts_close"); return -1; }
void *sam_hdr_parse(int l_text, const char *text) { (void)l_text; (void)text; return die("sam_hdr_parse"); }
int sam_hdr_write(void *fp, const void *h) { (void)fp; (void)h; die("sam_hdr_write"); return -1; }
int sam_parse1(void *s, void *h, void *b) { (void)s; (void)h; (void)b; die("sam_parse1"); return -1; }
int sam_write1(void *fp, const void *h, const void *b) { (void)fp; (void)h; (void)b; die("sam_write1"); return -1; }
