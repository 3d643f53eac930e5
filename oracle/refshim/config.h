/* oracle/refshim/config.h — TEST INFRASTRUCTURE ONLY. Empty build config for
 * compiling reference sources in place; ORC is disabled so the vendored C
 * backups (*-dist.c) are used. */
#ifndef B200_REFSHIM_CONFIG_H
#define B200_REFSHIM_CONFIG_H
#ifndef DISABLE_ORC
#define DISABLE_ORC 1
#endif
#define GST_API_VERSION "1.0"
#define PACKAGE "b200-refshim"
#define VERSION "1.29.2.1"
#endif
