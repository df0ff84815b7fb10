/* TEST INFRASTRUCTURE -- not part of the product.
 *
 * Stand-in for <libconfig.h> so that the reference's UNMODIFIED sources
 * (/root/reference/src/starrynight-main.c, which #includes <libconfig.h> at
 * line 16 and uses it only in starrynight-config.c:102-179) compile in an image
 * that has no libconfig.  Only the 13 entry points the reference calls exist.
 * It is backed by our own cfg parser (driver/sn_cfg.h), so the reference's
 * load_config() reads real starrynight.cfg files with libconfig's strict type
 * rules.
 */
#ifndef SN_STUB_LIBCONFIG_H
#define SN_STUB_LIBCONFIG_H

#include "../../driver/sn_cfg.h"

#define CONFIG_TRUE 1
#define CONFIG_FALSE 0

typedef snc_config config_t;
typedef snc_node config_setting_t;

static void config_init(config_t *c) { snc_init(c); }
static void config_destroy(config_t *c) { snc_destroy(c); }
static int config_read_file(config_t *c, const char *fn) { return snc_read_file(c, fn); }
static const char *config_error_file(const config_t *c) { return c->err_file; }
static int config_error_line(const config_t *c) { return c->err_line; }
static const char *config_error_text(const config_t *c) { return c->err_text; }
static int config_lookup_string(const config_t *c, const char *p, const char **v) { return snc_lookup_string(c, p, v); }
static int config_lookup_int(const config_t *c, const char *p, int *v) { return snc_lookup_int(c, p, v); }
static int config_lookup_float(const config_t *c, const char *p, double *v) { return snc_lookup_float(c, p, v); }
static int config_lookup_bool(const config_t *c, const char *p, int *v) { return snc_lookup_bool(c, p, v); }
static config_setting_t *config_lookup(const config_t *c, const char *p) { return (config_setting_t *)snc_lookup(c, p); }
static int config_setting_length(const config_setting_t *s) { return snc_length(s); }
static double config_setting_get_float_elem(const config_setting_t *s, int i) { return snc_get_float_elem(s, i); }

#endif
