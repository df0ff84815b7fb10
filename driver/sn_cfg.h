/* sn_cfg.h -- a small header-only parser for the subset of the libconfig
 * grammar that StarryNight's configuration files use.
 *
 * Why it exists: the reference reads ./starrynight.cfg through libconfig
 * (/root/reference/src/starrynight-config.c:98-182).  libconfig is not
 * available in this image, so the B200 driver (driver/starrynight_b200_main.c)
 * and the stub <libconfig.h> used to build the *unmodified* reference as a CPU
 * oracle (oracle/stub/libconfig.h) both sit on this parser.
 *
 * Grammar handled (libconfig manual, "Configuration Files"):
 *   setting   := name (':' | '=') value (';' | ',')?
 *   value     := scalar | '{' setting* '}' | '[' scalar,* ']' | '(' value,* ')'
 *   scalar    := int (dec / 0x hex, optional L/LL) | float | bool | "string"...
 *   comments  := '#' ..., '//' ..., C block comments
 * Type rules mirror libconfig's strict (non auto-convert) mode: a lookup of the
 * wrong scalar type fails and leaves the destination untouched.  That is what
 * makes "MCMoves: 200.0 #Must be floating point!" matter in the reference cfg
 * (/root/reference/starrynight.cfg:68).
 */
#ifndef SN_CFG_H
#define SN_CFG_H

#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef enum { SNC_NONE = 0, SNC_GROUP, SNC_INT, SNC_FLOAT, SNC_BOOL, SNC_STRING,
               SNC_ARRAY, SNC_LIST } snc_type;

typedef struct snc_node {
    char *name;              /* NULL for array / list elements */
    snc_type type;
    long long ival;          /* SNC_INT, SNC_BOOL */
    double fval;             /* SNC_FLOAT */
    char *sval;              /* SNC_STRING */
    struct snc_node **child; /* SNC_GROUP / SNC_ARRAY / SNC_LIST */
    int nchild, cap;
} snc_node;

typedef struct {
    snc_node *root;
    char err_text[160];
    int err_line;
    char err_file[256];
    /* parser state */
    const char *p;
    int line;
} snc_config;

static snc_node *snc_new(snc_type t, const char *name)
{
    snc_node *n = (snc_node *)calloc(1, sizeof(snc_node));
    n->type = t;
    if (name) { n->name = (char *)malloc(strlen(name) + 1); strcpy(n->name, name); }
    return n;
}

static void snc_add(snc_node *parent, snc_node *c)
{
    if (parent->nchild == parent->cap) {
        parent->cap = parent->cap ? parent->cap * 2 : 8;
        parent->child = (snc_node **)realloc(parent->child, sizeof(snc_node *) * parent->cap);
    }
    parent->child[parent->nchild++] = c;
}

static void snc_free_node(snc_node *n)
{
    int i;
    if (!n) return;
    for (i = 0; i < n->nchild; i++) snc_free_node(n->child[i]);
    free(n->child); free(n->name); free(n->sval); free(n);
}

static void snc_skip_ws(snc_config *c)
{
    for (;;) {
        while (*c->p && isspace((unsigned char)*c->p)) { if (*c->p == '\n') c->line++; c->p++; }
        if (*c->p == '#' || (c->p[0] == '/' && c->p[1] == '/')) {
            while (*c->p && *c->p != '\n') c->p++;
        } else if (c->p[0] == '/' && c->p[1] == '*') {
            c->p += 2;
            while (*c->p && !(c->p[0] == '*' && c->p[1] == '/')) { if (*c->p == '\n') c->line++; c->p++; }
            if (*c->p) c->p += 2;
        } else return;
    }
}

static int snc_fail(snc_config *c, const char *msg)
{
    if (!c->err_text[0]) { snprintf(c->err_text, sizeof c->err_text, "%s", msg); c->err_line = c->line; }
    return 0;
}

static int snc_parse_value(snc_config *c, snc_node *out);

static int snc_parse_scalar(snc_config *c, snc_node *out)
{
    const char *s = c->p;
    if (*s == '"') {                       /* string, adjacent strings concatenate */
        size_t len = 0, cap = 64; char *buf = (char *)malloc(cap);
        while (*c->p == '"') {
            c->p++;
            while (*c->p && *c->p != '"') {
                char ch = *c->p++;
                if (ch == '\\' && *c->p) {
                    char e = *c->p++;
                    ch = e == 'n' ? '\n' : e == 't' ? '\t' : e == 'r' ? '\r' : e == 'f' ? '\f' : e;
                }
                if (ch == '\n') c->line++;
                if (len + 2 > cap) { cap *= 2; buf = (char *)realloc(buf, cap); }
                buf[len++] = ch;
            }
            if (*c->p != '"') { free(buf); return snc_fail(c, "unterminated string"); }
            c->p++;
            snc_skip_ws(c);
        }
        buf[len] = 0; out->type = SNC_STRING; out->sval = buf; return 1;
    }
    if (isalpha((unsigned char)*s)) {      /* true / false, any case */
        size_t n = 0; while (isalpha((unsigned char)s[n])) n++;
        if (n == 4 && (s[0]=='t'||s[0]=='T') && (s[1]=='r'||s[1]=='R') && (s[2]=='u'||s[2]=='U') && (s[3]=='e'||s[3]=='E'))
        { out->type = SNC_BOOL; out->ival = 1; c->p += 4; return 1; }
        if (n == 5 && (s[0]=='f'||s[0]=='F') && (s[1]=='a'||s[1]=='A') && (s[2]=='l'||s[2]=='L') && (s[3]=='s'||s[3]=='S') && (s[4]=='e'||s[4]=='E'))
        { out->type = SNC_BOOL; out->ival = 0; c->p += 5; return 1; }
        return snc_fail(c, "syntax error");
    }
    {   /* number: float iff it has '.', or an exponent on a non-hex literal */
        const char *q = s; int is_float = 0, is_hex = 0; char *end;
        if (*q == '+' || *q == '-') q++;
        if (q[0] == '0' && (q[1] == 'x' || q[1] == 'X')) { is_hex = 1; q += 2; }
        if (!isdigit((unsigned char)*q) && !(*q == '.' && isdigit((unsigned char)q[1])) && !(is_hex && isxdigit((unsigned char)*q)))
            return snc_fail(c, "syntax error");
        while (*q && (isxdigit((unsigned char)*q) || *q == '.' || *q == '+' || *q == '-')) {
            if (*q == '.') is_float = 1;
            if (!is_hex && (*q == 'e' || *q == 'E')) is_float = 1;
            if ((*q == '+' || *q == '-') && !(q[-1] == 'e' || q[-1] == 'E')) break;
            if (!is_hex && isalpha((unsigned char)*q) && *q != 'e' && *q != 'E') break;
            q++;
        }
        if (is_float) { out->type = SNC_FLOAT; out->fval = strtod(s, &end); }
        else { out->type = SNC_INT; out->ival = strtoll(s, &end, 0); while (*end == 'L' || *end == 'l') end++; }
        if (end == s) return snc_fail(c, "syntax error");
        c->p = end; return 1;
    }
}

static int snc_parse_settings(snc_config *c, snc_node *group, int until_brace)
{
    for (;;) {
        char name[128]; size_t n = 0; snc_node *child;
        snc_skip_ws(c);
        if (!*c->p) return until_brace ? snc_fail(c, "missing '}'") : 1;
        if (*c->p == '}') { if (until_brace) { c->p++; return 1; } return snc_fail(c, "unexpected '}'"); }
        if (!(isalpha((unsigned char)*c->p) || *c->p == '*' || *c->p == '_')) return snc_fail(c, "syntax error");
        while ((isalnum((unsigned char)*c->p) || *c->p == '_' || *c->p == '-' || *c->p == '*') && n + 1 < sizeof name) name[n++] = *c->p++;
        name[n] = 0;
        snc_skip_ws(c);
        if (*c->p != ':' && *c->p != '=') return snc_fail(c, "syntax error");
        c->p++;
        snc_skip_ws(c);
        child = snc_new(SNC_NONE, name);
        if (!snc_parse_value(c, child)) { snc_free_node(child); return 0; }
        snc_add(group, child);
        snc_skip_ws(c);
        if (*c->p == ';' || *c->p == ',') c->p++;
    }
}

static int snc_parse_value(snc_config *c, snc_node *out)
{
    snc_skip_ws(c);
    if (*c->p == '{') { c->p++; out->type = SNC_GROUP; return snc_parse_settings(c, out, 1); }
    if (*c->p == '[' || *c->p == '(') {
        char close = *c->p == '[' ? ']' : ')';
        out->type = *c->p == '[' ? SNC_ARRAY : SNC_LIST;
        c->p++;
        for (;;) {
            snc_node *e;
            snc_skip_ws(c);
            if (*c->p == close) { c->p++; return 1; }
            if (!*c->p) return snc_fail(c, "unterminated array");
            e = snc_new(SNC_NONE, NULL);
            if (!(out->type == SNC_ARRAY ? snc_parse_scalar(c, e) : snc_parse_value(c, e))) { snc_free_node(e); return 0; }
            snc_add(out, e);
            snc_skip_ws(c);
            if (*c->p == ',') c->p++;
        }
    }
    return snc_parse_scalar(c, out);
}

static void snc_init(snc_config *c) { memset(c, 0, sizeof *c); }

static void snc_destroy(snc_config *c) { snc_free_node(c->root); c->root = NULL; }

/* returns 1 on success, 0 on failure (err_* filled in) */
static int snc_read_string(snc_config *c, const char *text)
{
    snc_free_node(c->root);
    c->root = snc_new(SNC_GROUP, NULL);
    c->p = text; c->line = 1; c->err_text[0] = 0;
    return snc_parse_settings(c, c->root, 0);
}

static int snc_read_file(snc_config *c, const char *path)
{
    FILE *f = fopen(path, "rb"); long sz; char *buf; int ok;
    snprintf(c->err_file, sizeof c->err_file, "%s", path);
    if (!f) { snprintf(c->err_text, sizeof c->err_text, "file I/O error"); c->err_line = 0; return 0; }
    fseek(f, 0, SEEK_END); sz = ftell(f); fseek(f, 0, SEEK_SET);
    buf = (char *)malloc((size_t)sz + 1);
    if (fread(buf, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(buf); snprintf(c->err_text, sizeof c->err_text, "file I/O error"); return 0; }
    buf[sz] = 0; fclose(f);
    ok = snc_read_string(c, buf);
    free(buf);
    return ok;
}

/* dotted-path lookup: "Efield.x" */
static const snc_node *snc_lookup(const snc_config *c, const char *path)
{
    const snc_node *cur = c->root;
    while (cur && *path) {
        size_t n = strcspn(path, "."); int i; const snc_node *next = NULL;
        if (cur->type != SNC_GROUP) return NULL;
        for (i = 0; i < cur->nchild; i++)
            if (cur->child[i]->name && strlen(cur->child[i]->name) == n && !strncmp(cur->child[i]->name, path, n)) { next = cur->child[i]; break; }
        cur = next; path += n; if (*path == '.') path++;
    }
    return cur;
}

static int snc_lookup_int(const snc_config *c, const char *path, int *v)
{ const snc_node *n = snc_lookup(c, path); if (!n || n->type != SNC_INT) return 0; *v = (int)n->ival; return 1; }
static int snc_lookup_float(const snc_config *c, const char *path, double *v)
{ const snc_node *n = snc_lookup(c, path); if (!n || n->type != SNC_FLOAT) return 0; *v = n->fval; return 1; }
static int snc_lookup_bool(const snc_config *c, const char *path, int *v)
{ const snc_node *n = snc_lookup(c, path); if (!n || n->type != SNC_BOOL) return 0; *v = (int)n->ival; return 1; }
static int snc_lookup_string(const snc_config *c, const char *path, const char **v)
{ const snc_node *n = snc_lookup(c, path); if (!n || n->type != SNC_STRING) return 0; *v = n->sval; return 1; }
static int snc_length(const snc_node *n)
{ return (n && (n->type == SNC_GROUP || n->type == SNC_ARRAY || n->type == SNC_LIST)) ? n->nchild : 0; }
/* strict: a non-float element yields 0.0, like config_setting_get_float_elem */
static double snc_get_float_elem(const snc_node *n, int i)
{ if (!n || i < 0 || i >= snc_length(n) || n->child[i]->type != SNC_FLOAT) return 0.0; return n->child[i]->fval; }
static long long snc_get_int_elem(const snc_node *n, int i)
{ if (!n || i < 0 || i >= snc_length(n) || n->child[i]->type != SNC_INT) return 0; return n->child[i]->ival; }

#endif /* SN_CFG_H */
