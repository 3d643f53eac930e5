/* oracle/refshim/glib.h — TEST INFRASTRUCTURE ONLY.
 *
 * A minimal stand-in for <glib.h> so that a handful of the reference's own
 * arithmetic source files (video-converter.c, video-scaler.c, blend.c,
 * audio-resampler.c, the ORC "-dist.c" C backups ...) can be compiled
 * *in place* from /root/reference into oracle/_ref/libgstref.so without a
 * GLib installation (GLib is a network meson wrap, absent in this image).
 *
 * Nothing here is product code; nothing here is copied from GLib.  It only
 * provides the type names / allocation macros / containers those files use.
 */
#ifndef B200_REFSHIM_GLIB_H
#define B200_REFSHIM_GLIB_H

#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <stdarg.h>
#include <limits.h>
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <alloca.h>

#ifdef __cplusplus
#define G_BEGIN_DECLS extern "C" {
#define G_END_DECLS }
#else
#define G_BEGIN_DECLS
#define G_END_DECLS
#endif

G_BEGIN_DECLS

typedef char gchar;
typedef unsigned char guchar;
typedef short gshort;
typedef unsigned short gushort;
typedef int gint;
typedef unsigned int guint;
typedef long glong;
typedef unsigned long gulong;
typedef int8_t gint8;
typedef uint8_t guint8;
typedef int16_t gint16;
typedef uint16_t guint16;
typedef int32_t gint32;
typedef uint32_t guint32;
typedef int64_t gint64;
typedef uint64_t guint64;
typedef float gfloat;
typedef double gdouble;
typedef int gboolean;
typedef void *gpointer;
typedef const void *gconstpointer;
typedef size_t gsize;
typedef ptrdiff_t gssize;
typedef intptr_t gintptr;
typedef uintptr_t guintptr;
typedef unsigned long GType;
typedef guint32 GQuark;
typedef void (*GDestroyNotify) (gpointer data);
typedef void (*GFunc) (gpointer data, gpointer user_data);
typedef gint (*GCompareFunc) (gconstpointer a, gconstpointer b);
typedef gpointer (*GBoxedCopyFunc) (gpointer boxed);
typedef void (*GBoxedFreeFunc) (gpointer boxed);

#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif
#ifndef NULL
#define NULL ((void*)0)
#endif

#undef MIN
#undef MAX
#undef ABS
#undef CLAMP
#define MIN(a,b) (((a) < (b)) ? (a) : (b))
#define MAX(a,b) (((a) > (b)) ? (a) : (b))
#define ABS(a) (((a) < 0) ? -(a) : (a))
#define CLAMP(x,lo,hi) (((x) > (hi)) ? (hi) : (((x) < (lo)) ? (lo) : (x)))

#define G_N_ELEMENTS(a) (sizeof (a) / sizeof ((a)[0]))
#define G_STRINGIFY(x) G_STRINGIFY_ARG(x)
#define G_STRINGIFY_ARG(x) #x
#define G_STRLOC __FILE__
#define G_STRFUNC ((const char*) (__func__))
#define G_GNUC_CONST __attribute__((__const__))
#define G_GNUC_PURE __attribute__((__pure__))
#define G_GNUC_MALLOC __attribute__((__malloc__))
#define G_GNUC_UNUSED __attribute__((__unused__))
#define G_GNUC_INTERNAL __attribute__((visibility("hidden")))
#define G_GNUC_WARN_UNUSED_RESULT
#define G_GNUC_NULL_TERMINATED
#define G_GNUC_PRINTF(a,b)
#define G_GNUC_NO_INSTRUMENT
#define G_GNUC_BEGIN_IGNORE_DEPRECATIONS
#define G_GNUC_END_IGNORE_DEPRECATIONS
#define G_GNUC_DEPRECATED
#define G_GNUC_DEPRECATED_FOR(f)
#define G_DEPRECATED
#define G_DEPRECATED_FOR(f)
#define G_UNAVAILABLE(a,b)
#define G_LIKELY(x) (__builtin_expect (!!(x), 1))
#define G_UNLIKELY(x) (__builtin_expect (!!(x), 0))
#define G_STATIC_ASSERT(e) _Static_assert (e, "static assert")
#define G_INLINE_FUNC static inline
#define G_ALWAYS_INLINE __attribute__((always_inline))
#define G_NO_INLINE __attribute__((noinline))
#define G_STMT_START do
#define G_STMT_END while (0)
#define G_DEFINE_AUTOPTR_CLEANUP_FUNC(t,f)
#define G_DEFINE_AUTO_CLEANUP_CLEAR_FUNC(t,f)
#define G_DEFINE_AUTO_CLEANUP_FREE_FUNC(t,f,n)
#define G_GSIZE_FORMAT "lu"
#define G_GSSIZE_FORMAT "ld"
#define G_GINT64_FORMAT "ld"
#define G_GUINT64_FORMAT "lu"
#define G_GINT64_MODIFIER "l"
#define G_GINT64_CONSTANT(v) (v##L)
#define G_GUINT64_CONSTANT(v) (v##UL)

#define G_MININT INT_MIN
#define G_MAXINT INT_MAX
#define G_MAXUINT UINT_MAX
#define G_MININT8 ((gint8) -128)
#define G_MAXINT8 ((gint8) 127)
#define G_MAXUINT8 ((guint8) 255)
#define G_MININT16 ((gint16) -32768)
#define G_MAXINT16 ((gint16) 32767)
#define G_MAXUINT16 ((guint16) 65535)
#define G_MININT32 ((gint32) (-2147483647 - 1))
#define G_MAXINT32 ((gint32) 2147483647)
#define G_MAXUINT32 ((guint32) 0xffffffffU)
#define G_MININT64 INT64_MIN
#define G_MAXINT64 INT64_MAX
#define G_MAXUINT64 UINT64_MAX
#define G_MAXSIZE SIZE_MAX
#define G_MAXDOUBLE DBL_MAX
#define G_MINDOUBLE DBL_MIN
#define G_MAXFLOAT FLT_MAX
#define G_PI 3.1415926535897932384626433832795028841971693993751
#define G_PI_2 1.5707963267948966192313216916397514420985846996876

#define G_LITTLE_ENDIAN 1234
#define G_BIG_ENDIAN 4321
#define G_BYTE_ORDER G_LITTLE_ENDIAN

#define GINT_TO_POINTER(i) ((gpointer) (glong) (i))
#define GPOINTER_TO_INT(p) ((gint) (glong) (p))
#define GUINT_TO_POINTER(u) ((gpointer) (gulong) (u))
#define GPOINTER_TO_UINT(p) ((guint) (gulong) (p))
#define GSIZE_TO_POINTER(s) ((gpointer) (gsize) (s))
#define GPOINTER_TO_SIZE(p) ((gsize) (p))

#define GUINT16_SWAP_LE_BE(v) ((guint16) __builtin_bswap16 ((guint16) (v)))
#define GUINT32_SWAP_LE_BE(v) ((guint32) __builtin_bswap32 ((guint32) (v)))
#define GUINT64_SWAP_LE_BE(v) ((guint64) __builtin_bswap64 ((guint64) (v)))
#define GUINT16_FROM_LE(v) ((guint16)(v))
#define GUINT16_TO_LE(v) ((guint16)(v))
#define GUINT32_FROM_LE(v) ((guint32)(v))
#define GUINT32_TO_LE(v) ((guint32)(v))
#define GUINT16_FROM_BE(v) GUINT16_SWAP_LE_BE(v)
#define GUINT16_TO_BE(v) GUINT16_SWAP_LE_BE(v)
#define GUINT32_FROM_BE(v) GUINT32_SWAP_LE_BE(v)
#define GUINT32_TO_BE(v) GUINT32_SWAP_LE_BE(v)
#define GUINT64_FROM_LE(v) ((guint64)(v))
#define GUINT64_TO_LE(v) ((guint64)(v))
#define GUINT64_FROM_BE(v) GUINT64_SWAP_LE_BE(v)
#define GUINT64_TO_BE(v) GUINT64_SWAP_LE_BE(v)
#define GINT16_FROM_BE(v) ((gint16) GUINT16_SWAP_LE_BE(v))
#define GINT16_TO_BE(v) ((gint16) GUINT16_SWAP_LE_BE(v))
#define GINT32_FROM_BE(v) ((gint32) GUINT32_SWAP_LE_BE(v))
#define GINT32_TO_BE(v) ((gint32) GUINT32_SWAP_LE_BE(v))

/* ---- diagnostics: the reference's argument checks stay live ------------- */
void refshim_critical (const char *file, int line, const char *expr);
#define g_return_if_fail(e) do { if (!(e)) { refshim_critical (__FILE__, __LINE__, #e); return; } } while (0)
#define g_return_val_if_fail(e,v) do { if (!(e)) { refshim_critical (__FILE__, __LINE__, #e); return (v); } } while (0)
#define g_return_if_reached() do { refshim_critical (__FILE__, __LINE__, "reached"); return; } while (0)
#define g_return_val_if_reached(v) do { refshim_critical (__FILE__, __LINE__, "reached"); return (v); } while (0)
#define g_assert(e) do { if (!(e)) { refshim_critical (__FILE__, __LINE__, #e); abort (); } } while (0)
#define g_assert_not_reached() do { refshim_critical (__FILE__, __LINE__, "not reached"); abort (); } while (0)
#define g_assert_cmpint(a,op,b) g_assert ((a) op (b))
#define g_assert_cmpuint(a,op,b) g_assert ((a) op (b))
#define g_warning(...) do { fprintf (stderr, "refshim warning: " __VA_ARGS__); fputc ('\n', stderr); } while (0)
#define g_critical(...) do { fprintf (stderr, "refshim critical: " __VA_ARGS__); fputc ('\n', stderr); } while (0)
#define g_message(...) do { } while (0)
#define g_debug(...) do { } while (0)
#define g_error(...) do { fprintf (stderr, "refshim error: " __VA_ARGS__); abort (); } while (0)
#define g_print printf
#define g_printerr(...) fprintf (stderr, __VA_ARGS__)

/* ---- memory --------------------------------------------------------------- */
static inline gpointer g_malloc (gsize n) { return n ? malloc (n) : NULL; }
static inline gpointer g_malloc0 (gsize n) { return n ? calloc (1, n) : NULL; }
static inline gpointer g_realloc (gpointer p, gsize n) { if (!n) { free (p); return NULL; } return realloc (p, n); }
static inline void g_free (gpointer p) { free (p); }
static inline gpointer g_memdup2 (gconstpointer p, gsize n) { gpointer r; if (!p || !n) return NULL; r = malloc (n); memcpy (r, p, n); return r; }
#define g_memdup(p,n) g_memdup2 (p, n)
static inline gpointer g_realloc_n (gpointer p, gsize n, gsize sz) { return g_realloc (p, n * sz); }
static inline gpointer g_malloc_n (gsize n, gsize sz) { return g_malloc (n * sz); }
static inline gpointer g_malloc0_n (gsize n, gsize sz) { return g_malloc0 (n * sz); }
#define g_new(t,n) ((t *) g_malloc (sizeof (t) * (gsize) (n)))
#define g_new0(t,n) ((t *) g_malloc0 (sizeof (t) * (gsize) (n)))
#define g_renew(t,p,n) ((t *) g_realloc (p, sizeof (t) * (gsize) (n)))
#define g_alloca(n) alloca (n)
#define g_newa(t,n) ((t *) alloca (sizeof (t) * (gsize) (n)))
#define g_slice_new(t) g_new (t, 1)
#define g_slice_new0(t) g_new0 (t, 1)
#define g_slice_free(t,p) g_free (p)
#define g_clear_pointer(pp,destroy) do { if (*(pp)) { destroy (*(pp)); *(pp) = NULL; } } while (0)
#define g_steal_pointer(pp) refshim_steal_pointer ((gpointer) (pp))
static inline gpointer refshim_steal_pointer (gpointer pp) { gpointer *ptr = (gpointer *) pp; gpointer r = *ptr; *ptr = NULL; return r; }

/* ---- strings -------------------------------------------------------------- */
static inline gboolean g_str_equal (gconstpointer a, gconstpointer b) { return strcmp ((const char *) a, (const char *) b) == 0; }
static inline gboolean g_str_has_prefix (const gchar * s, const gchar * p) { return strncmp (s, p, strlen (p)) == 0; }
static inline gchar *g_strdup (const gchar * s) { return s ? strdup (s) : NULL; }
gchar *g_strdup_printf (const gchar * fmt, ...);
static inline gint g_strcmp0 (const char *a, const char *b) { if (!a) return -(a != b); if (!b) return a != b; return strcmp (a, b); }
#define g_ascii_strcasecmp strcasecmp
#define g_snprintf snprintf

typedef struct { gchar *str; gsize len; gsize allocated_len; } GString;
GString *g_string_new (const gchar * init);
GString *g_string_append (GString * s, const gchar * v);
gchar *g_string_free (GString * s, gboolean free_segment);
gchar **g_strsplit (const gchar * s, const gchar * delim, gint max);
void g_strfreev (gchar ** v);

/* ---- once / threads --------------------------------------------------------- */
static inline gboolean g_once_init_enter (volatile void *loc) { return *(volatile gsize *) loc == 0; }
static inline void g_once_init_leave (volatile void *loc, gsize v) { *(volatile gsize *) loc = v; }
typedef struct { volatile int status; volatile gpointer retval; } GOnce;
#define G_ONCE_INIT { 0, NULL }
typedef gpointer (*GThreadFunc) (gpointer data);
gpointer refshim_once (GOnce * once, GThreadFunc func, gpointer arg);
#define g_once(once,func,arg) refshim_once ((once), (func), (arg))

typedef struct { pthread_mutex_t m; int inited; } GMutex;
typedef struct { pthread_cond_t c; int inited; } GCond;
static inline void g_mutex_init (GMutex * m) { pthread_mutex_init (&m->m, NULL); m->inited = 1; }
static inline void g_mutex_clear (GMutex * m) { pthread_mutex_destroy (&m->m); }
static inline void g_mutex_lock (GMutex * m) { pthread_mutex_lock (&m->m); }
static inline void g_mutex_unlock (GMutex * m) { pthread_mutex_unlock (&m->m); }
static inline void g_cond_init (GCond * c) { pthread_cond_init (&c->c, NULL); }
static inline void g_cond_clear (GCond * c) { pthread_cond_destroy (&c->c); }
static inline void g_cond_wait (GCond * c, GMutex * m) { pthread_cond_wait (&c->c, &m->m); }
static inline void g_cond_signal (GCond * c) { pthread_cond_signal (&c->c); }
static inline void g_cond_broadcast (GCond * c) { pthread_cond_broadcast (&c->c); }
guint g_get_num_processors (void);
#define g_atomic_int_get(p) __atomic_load_n ((p), __ATOMIC_SEQ_CST)
#define g_atomic_int_set(p,v) __atomic_store_n ((p), (v), __ATOMIC_SEQ_CST)
#define g_atomic_int_inc(p) ((void) __atomic_fetch_add ((p), 1, __ATOMIC_SEQ_CST))
#define g_atomic_int_add(p,v) __atomic_fetch_add ((p), (v), __ATOMIC_SEQ_CST)
#define g_atomic_int_dec_and_test(p) (__atomic_fetch_sub ((p), 1, __ATOMIC_SEQ_CST) == 1)

/* ---- GPtrArray ---------------------------------------------------------------- */
typedef struct { gpointer *pdata; guint len; guint alloc; } GPtrArray;
GPtrArray *g_ptr_array_new (void);
void g_ptr_array_unref (GPtrArray * a);
void g_ptr_array_set_size (GPtrArray * a, gint len);
void g_ptr_array_add (GPtrArray * a, gpointer p);
void g_ptr_array_remove_range (GPtrArray * a, guint index, guint len);

/* ---- GType / GValue stubs (never exercised on the arithmetic paths) ---------- */
#define G_TYPE_INVALID ((GType) 0)
#define G_TYPE_BOOLEAN ((GType) 20)
#define G_TYPE_INT ((GType) 24)
#define G_TYPE_UINT ((GType) 28)
#define G_TYPE_INT64 ((GType) 40)
#define G_TYPE_UINT64 ((GType) 44)
#define G_TYPE_ENUM ((GType) 48)
#define G_TYPE_FLAGS ((GType) 52)
#define G_TYPE_FLOAT ((GType) 56)
#define G_TYPE_DOUBLE ((GType) 60)
#define G_TYPE_STRING ((GType) 64)
#define G_TYPE_POINTER ((GType) 68)
#define REFSHIM_TYPE_ENUM_BASE ((GType) 1000)   /* every GST_TYPE_*enum* maps at/above this */

typedef struct { GType g_type; union { gint i; guint u; gint64 i64; gdouble d; gpointer p; } data[2]; } GValue;
#define G_VALUE_INIT { 0, { { 0 } } }
GValue *g_value_init (GValue * v, GType t);
void g_value_unset (GValue * v);
void g_value_set_static_string (GValue * v, const gchar * s);
const gchar *g_value_get_string (const GValue * v);
#define G_VALUE_TYPE(v) ((v)->g_type)
#define G_VALUE_HOLDS(v,t) ((v)->g_type == (t))

typedef struct { guint value; const gchar *value_name; const gchar *value_nick; } GFlagsValue;
typedef struct { gint value; const gchar *value_name; const gchar *value_nick; } GEnumValue;
typedef struct { GType g_type; guint mask; guint n_values; GFlagsValue *values; } GFlagsClass;
typedef struct { GType g_type; gint minimum, maximum; guint n_values; GEnumValue *values; } GEnumClass;
gpointer g_type_class_ref (GType t);
void g_type_class_unref (gpointer k);
GFlagsValue *g_flags_get_value_by_nick (GFlagsClass * k, const gchar * nick);
GFlagsValue *g_flags_get_first_value (GFlagsClass * k, guint value);

#define G_DEFINE_BOXED_TYPE(TypeName,type_name,copy,free) \
  GType type_name##_get_type (void) { (void) copy; (void) free; return (GType) 0; }
#define G_DEFINE_POINTER_TYPE(TypeName,type_name) \
  GType type_name##_get_type (void) { return (GType) 0; }

typedef struct _GObject { int dummy; } GObject;
static inline gpointer g_object_ref (gpointer o) { return o; }
static inline void g_object_unref (gpointer o) { (void) o; }
#define G_OBJECT(o) ((GObject *) (o))

G_END_DECLS
#endif /* B200_REFSHIM_GLIB_H */
