/*
 * b9_dump.h — header-only writer of the golden-vector dump format (tests/golden_io.py).
 * TEST INFRASTRUCTURE: for the instrumented harness that will call the reference's own
 * likelihood once base-cpp is staged, and for the oracle's own self-checks.  It knows
 * nothing about the reference; it only writes doubles as C99 hex floats ("%a"), which
 * survive a text round trip bit for bit.  C99 / C++11, no dependencies.
 *
 *     b9dump_t d;
 *     b9dump_open(&d, "cfg1_logpost.b9dump");
 *     b9dump_meta(&d, "commit", "<sha>");  b9dump_meta(&d, "seed", "42");
 *     b9dump_record(&d, "stage2_mags", star, mags, n_bands);     // once per (stage, star)
 *     b9dump_record(&d, "logpost", -1, &lp, 1);                  // star -1: per-cluster value
 *     b9dump_close(&d);                                          // writes "end <count>"
 *
 * Every function returns 0 on success and -1 on an I/O or argument error.
 */
#ifndef B9_DUMP_H
#define B9_DUMP_H

#include <math.h>
#include <stdio.h>
#include <string.h>

typedef struct {
    FILE *f;
    long n_records;
} b9dump_t;

static int b9dump__name_ok(const char *s) {
    if (!s || !*s) return 0;
    for (; *s; ++s) {
        const char c = *s;
        if (!((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') ||
              c == '_' || c == '.' || c == '-'))
            return 0;
    }
    return 1;
}

static inline int b9dump_open(b9dump_t *d, const char *path) {
    if (!d || !path) return -1;
    d->n_records = 0;
    d->f = fopen(path, "w");
    if (!d->f) return -1;
    return fprintf(d->f, "b9dump 1\n") < 0 ? -1 : 0;
}

static inline int b9dump_meta(b9dump_t *d, const char *key, const char *value) {
    if (!d || !d->f || !b9dump__name_ok(key) || !value || strchr(value, '\n') || d->n_records) return -1;
    return fprintf(d->f, "meta %s %s\n", key, value) < 0 ? -1 : 0;
}

static inline int b9dump_record(b9dump_t *d, const char *stage, long star, const double *v, long n) {
    if (!d || !d->f || !b9dump__name_ok(stage) || star < -1 || n < 0 || (n > 0 && !v)) return -1;
    if (fprintf(d->f, "rec %s %ld %ld\n", stage, star, n) < 0) return -1;
    for (long i = 0; i < n; ++i) {
        int rc;
        if (isnan(v[i])) rc = fprintf(d->f, "nan");              /* sign/payload of a NaN is not data */
        else if (isinf(v[i])) rc = fprintf(d->f, v[i] > 0 ? "inf" : "-inf");
        else rc = fprintf(d->f, "%a", v[i]);
        if (rc < 0 || fputc((i % 4 == 3 || i == n - 1) ? '\n' : ' ', d->f) == EOF) return -1;
    }
    d->n_records++;
    return 0;
}

static inline int b9dump_close(b9dump_t *d) {
    if (!d || !d->f) return -1;
    int rc = fprintf(d->f, "end %ld\n", d->n_records) < 0 ? -1 : 0;
    if (fclose(d->f) != 0) rc = -1;
    d->f = NULL;
    return rc;
}

#endif /* B9_DUMP_H */
