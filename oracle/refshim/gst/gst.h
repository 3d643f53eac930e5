/* oracle/refshim/gst/gst.h — TEST INFRASTRUCTURE ONLY (see ../glib.h).
 *
 * Stand-in for <gst/gst.h>: just enough of the core API surface (GstStructure
 * option bags, GstTaskPool, byte read/write macros, rounding helpers, no-op
 * logging) for the reference's arithmetic sources to compile unmodified.
 * Written from scratch; not a copy of GStreamer core.
 */
#ifndef B200_REFSHIM_GST_H
#define B200_REFSHIM_GST_H

#include <glib.h>

G_BEGIN_DECLS

#define GST_API_EXPORT extern
#define GST_API_IMPORT extern
#define GST_API extern
#define GST_PADDING 4
#define GST_PADDING_LARGE 20

/* ---- logging: compiled out ------------------------------------------------- */
#ifndef GST_DISABLE_GST_DEBUG
#define GST_DISABLE_GST_DEBUG 1
#endif
typedef struct _GstDebugCategory GstDebugCategory;
#define GST_DEBUG_CATEGORY_STATIC(c)
#define GST_DEBUG_CATEGORY_EXTERN(c)
#define GST_DEBUG_CATEGORY(c)
#define GST_DEBUG_CATEGORY_INIT(...) do { } while (0)
#define GST_DEBUG_CATEGORY_GET(c,n) do { } while (0)
#define GST_LOG(...) do { } while (0)
#define GST_DEBUG(...) do { } while (0)
#define GST_INFO(...) do { } while (0)
#define GST_TRACE(...) do { } while (0)
#define GST_FIXME(...) do { } while (0)
#define GST_WARNING(...) do { } while (0)
#define GST_ERROR(...) do { } while (0)
#define GST_LOG_OBJECT(...) do { } while (0)
#define GST_DEBUG_OBJECT(...) do { } while (0)
#define GST_INFO_OBJECT(...) do { } while (0)
#define GST_WARNING_OBJECT(...) do { } while (0)
#define GST_ERROR_OBJECT(...) do { } while (0)
#define GST_CAT_DEBUG(...) do { } while (0)
#define GST_CAT_LOG(...) do { } while (0)
#define GST_CAT_DEBUG_OBJECT(...) do { } while (0)
#define GST_PTR_FORMAT "p"
#define GST_DEBUG_FUNCPTR(f) (f)

/* ---- misc scalar helpers ------------------------------------------------------ */
typedef guint64 GstClockTime;
typedef gint64 GstClockTimeDiff;
#define GST_SECOND ((GstClockTime) 1000000000)
#define GST_MSECOND ((GstClockTime) 1000000)
#define GST_CLOCK_TIME_NONE ((GstClockTime) -1)
#define GST_CLOCK_TIME_IS_VALID(t) (((GstClockTime) (t)) != GST_CLOCK_TIME_NONE)
typedef enum { GST_FORMAT_UNDEFINED = 0, GST_FORMAT_DEFAULT = 1, GST_FORMAT_BYTES = 2,
  GST_FORMAT_TIME = 3, GST_FORMAT_BUFFERS = 4, GST_FORMAT_PERCENT = 5 } GstFormat;
const gchar *gst_format_get_name (GstFormat f);
guint64 gst_util_uint64_scale (guint64 val, guint64 num, guint64 denom);
guint64 gst_util_uint64_scale_int (guint64 val, gint num, gint denom);
gint gst_util_greatest_common_divisor (gint a, gint b);

#define GST_ROUND_UP_2(n) (((n) + 1) & ~1)
#define GST_ROUND_UP_4(n) (((n) + 3) & ~3)
#define GST_ROUND_UP_8(n) (((n) + 7) & ~7)
#define GST_ROUND_UP_16(n) (((n) + 15) & ~15)
#define GST_ROUND_UP_32(n) (((n) + 31) & ~31)
#define GST_ROUND_UP_64(n) (((n) + 63) & ~63)
#define GST_ROUND_UP_128(n) (((n) + 127) & ~127)
#define GST_ROUND_UP_N(n,a) ((((n) + ((a) - 1)) & ~((a) - 1)))
#define GST_ROUND_DOWN_2(n) ((n) & ~1)
#define GST_ROUND_DOWN_4(n) ((n) & ~3)
#define GST_ROUND_DOWN_8(n) ((n) & ~7)
#define GST_ROUND_DOWN_16(n) ((n) & ~15)
#define GST_ROUND_DOWN_N(n,a) ((n) & ~((a) - 1))
#define GST_MAKE_FOURCC(a,b,c,d) ((guint32) ((a) | (b) << 8 | (c) << 16 | (d) << 24))
#define GST_STR_NULL(s) ((s) ? (s) : "(NULL)")

static inline guint8 refshim_rd8 (const void *p) { return *(const guint8 *) p; }
static inline guint16 refshim_rd16 (const void *p) { guint16 v; memcpy (&v, p, 2); return v; }
static inline guint32 refshim_rd32 (const void *p) { guint32 v; memcpy (&v, p, 4); return v; }
static inline guint64 refshim_rd64 (const void *p) { guint64 v; memcpy (&v, p, 8); return v; }
static inline void refshim_wr16 (void *p, guint16 v) { memcpy (p, &v, 2); }
static inline void refshim_wr32 (void *p, guint32 v) { memcpy (p, &v, 4); }
static inline void refshim_wr64 (void *p, guint64 v) { memcpy (p, &v, 8); }
#define GST_READ_UINT8(p) refshim_rd8 (p)
#define GST_READ_UINT16_LE(p) refshim_rd16 (p)
#define GST_READ_UINT16_BE(p) GUINT16_SWAP_LE_BE (refshim_rd16 (p))
#define GST_READ_UINT32_LE(p) refshim_rd32 (p)
#define GST_READ_UINT32_BE(p) GUINT32_SWAP_LE_BE (refshim_rd32 (p))
#define GST_READ_UINT64_LE(p) refshim_rd64 (p)
#define GST_READ_UINT64_BE(p) GUINT64_SWAP_LE_BE (refshim_rd64 (p))
#define GST_WRITE_UINT8(p,v) (*(guint8 *) (p) = (guint8) (v))
#define GST_WRITE_UINT16_LE(p,v) refshim_wr16 ((p), (guint16) (v))
#define GST_WRITE_UINT16_BE(p,v) refshim_wr16 ((p), GUINT16_SWAP_LE_BE ((guint16) (v)))
#define GST_WRITE_UINT32_LE(p,v) refshim_wr32 ((p), (guint32) (v))
#define GST_WRITE_UINT32_BE(p,v) refshim_wr32 ((p), GUINT32_SWAP_LE_BE ((guint32) (v)))
#define GST_WRITE_UINT64_LE(p,v) refshim_wr64 ((p), (guint64) (v))
#define GST_WRITE_UINT64_BE(p,v) refshim_wr64 ((p), GUINT64_SWAP_LE_BE ((guint64) (v)))
static inline gfloat refshim_rdf (const void *p) { gfloat v; memcpy (&v, p, 4); return v; }
static inline void refshim_wrf (void *p, gfloat v) { memcpy (p, &v, 4); }
#define GST_READ_FLOAT_LE(p) refshim_rdf (p)
#define GST_WRITE_FLOAT_LE(p,v) refshim_wrf ((p), (v))

/* ---- opaque / skeletal core objects --------------------------------------------- */
typedef struct _GstObject { int refcount; } GstObject;
typedef struct _GstMiniObject { int refcount; guint flags; } GstMiniObject;
typedef struct _GstCaps GstCaps;
typedef struct _GstCapsFeatures GstCapsFeatures;
typedef struct _GstBufferPool GstBufferPool;
typedef struct _GstAllocator GstAllocator;
typedef struct _GstMemory GstMemory;
typedef struct _GstBuffer { GstMiniObject mini_object; } GstBuffer;
typedef struct _GstMetaInfo GstMetaInfo;
typedef struct _GstMeta { guint flags; const GstMetaInfo *info; } GstMeta;
typedef gboolean (*GstMetaTransformFunction) (GstBuffer * transbuf, GstMeta * meta,
    GstBuffer * buffer, GQuark type, gpointer data);
struct _GstMetaInfo { GType api; GType type; gsize size; gpointer init_func, free_func;
  GstMetaTransformFunction transform_func; };
typedef struct { gsize align; gsize prefix; gsize padding; guint flags; } GstAllocationParams;

typedef enum { GST_MAP_READ = 1, GST_MAP_WRITE = 2, GST_MAP_FLAG_LAST = (1 << 16) } GstMapFlags;
#define GST_MAP_READWRITE ((GstMapFlags) (GST_MAP_READ | GST_MAP_WRITE))
typedef struct { GstMemory *memory; GstMapFlags flags; guint8 *data; gsize size; gsize maxsize;
  gpointer user_data[4]; gpointer _gst_reserved[GST_PADDING]; } GstMapInfo;

#define GST_MINI_OBJECT_FLAG_LAST (1 << 4)
#define GST_BUFFER_FLAG_MARKER (GST_MINI_OBJECT_FLAG_LAST << 5)
#define GST_BUFFER_FLAG_LAST (GST_MINI_OBJECT_FLAG_LAST << 16)
#define GST_BUFFER_FLAGS(b) (((GstBuffer *) (b))->mini_object.flags)
#define GST_BUFFER_FLAG_IS_SET(b,f) (!!(GST_BUFFER_FLAGS (b) & (f)))
#define GST_CAPS_FEATURE_MEMORY_SYSTEM_MEMORY "memory:SystemMemory"
gboolean gst_buffer_is_writable (GstBuffer * b);
GstMeta *gst_buffer_iterate_meta (GstBuffer * b, gpointer * state);
gboolean gst_meta_api_type_tags_contain_only (GType api, const gchar ** tags);
static inline gpointer gst_object_ref (gpointer o) { return o; }
static inline void gst_object_unref (gpointer o) { (void) o; }

typedef struct _GstStructure GstStructure;
typedef struct _GstElement GstElement;
typedef struct _GstPad GstPad;
typedef struct _GstQuery GstQuery;
typedef struct _GstEvent GstEvent;
typedef struct _GstVideoCodecState GstVideoCodecState;
typedef struct _GstVideoCodecFrame GstVideoCodecFrame;

/* caps / GValue-list API: declared so the reference's caps helpers compile;
 * the shim implementations abort() because nothing on the pixel path calls them */
gboolean gst_value_deserialize (GValue * dest, const gchar * src);
guint gst_value_list_get_size (const GValue * v);
const GValue *gst_value_list_get_value (const GValue * v, guint i);
void gst_value_list_append_and_take_value (GValue * v, GValue * a);
gboolean gst_caps_is_fixed (const GstCaps * c);
GstCapsFeatures *gst_caps_get_features (const GstCaps * c, guint i);
gboolean gst_caps_features_contains (const GstCapsFeatures * f, const gchar * s);
GstCaps *gst_caps_new_static_str_simple (const gchar * media, const gchar * field, ...);
void gst_caps_set_simple_static_str (GstCaps * c, const gchar * field, ...);
void gst_caps_set_simple (GstCaps * c, const gchar * field, ...);
GstCapsFeatures *gst_caps_features_new_static_str (const gchar * f, ...);
void gst_caps_set_features (GstCaps * c, guint i, GstCapsFeatures * f);
GstCaps *gst_caps_new_full (gpointer s, ...);

/* ---- GstStructure: a tiny typed key/value bag ------------------------------------ */
typedef struct _GstIdStr { const gchar *s; } GstIdStr;
typedef gboolean (*GstStructureForeachIdStrFunc) (const GstIdStr * fieldname,
    const GValue * value, gpointer user_data);
GstStructure *gst_structure_new_empty (const gchar * name);
GstStructure *gst_structure_new_static_str_empty (const gchar * name);
GstStructure *gst_structure_new (const gchar * name, const gchar * firstfield, ...);
GstStructure *gst_structure_copy (const GstStructure * s);
void gst_structure_free (GstStructure * s);
void gst_structure_set (GstStructure * s, const gchar * field, ...);
void gst_structure_set_static_str (GstStructure * s, const gchar * field, ...);
void gst_structure_id_str_set_value (GstStructure * s, const GstIdStr * f, const GValue * v);
gboolean gst_structure_foreach_id_str (const GstStructure * s,
    GstStructureForeachIdStrFunc func, gpointer user_data);
gboolean gst_structure_get_int (const GstStructure * s, const gchar * f, gint * v);
gboolean gst_structure_get_uint (const GstStructure * s, const gchar * f, guint * v);
gboolean gst_structure_get_double (const GstStructure * s, const gchar * f, gdouble * v);
gboolean gst_structure_get_boolean (const GstStructure * s, const gchar * f, gboolean * v);
gboolean gst_structure_get_enum (const GstStructure * s, const gchar * f, GType t, gint * v);
const gchar *gst_structure_get_string (const GstStructure * s, const gchar * f);
const gchar *gst_structure_get_name (const GstStructure * s);
gboolean gst_structure_has_name (const GstStructure * s, const gchar * n);
gboolean gst_structure_has_field (const GstStructure * s, const gchar * f);
gboolean gst_structure_get_fraction (const GstStructure * s, const gchar * f, gint * n, gint * d);
GstStructure *gst_structure_new_static_str (const gchar * name, const gchar * firstfield, ...);
void gst_structure_take_value_static_str (GstStructure * s, const gchar * f, GValue * v);
GstStructure *gst_caps_get_structure (const GstCaps * c, guint i);
gboolean gst_structure_get_flagset (const GstStructure * s, const gchar * f, guint * fl, guint * m);
#define GST_TYPE_FRACTION ((GType) 200)
#define GST_TYPE_LIST ((GType) 204)
#define GST_TYPE_INT_RANGE ((GType) 208)
#define GST_TYPE_FRACTION_RANGE ((GType) 212)
#define GST_FLAG_SET_MASK_EXACT ((guint) -1)

/* ---- task pool (pthread-per-push) -------------------------------------------------- */
typedef struct _GstTaskPool { GstObject object; guint max_threads; } GstTaskPool;
typedef GstTaskPool GstSharedTaskPool;
typedef void (*GstTaskPoolFunction) (void *user_data);
typedef struct _GError GError;
#define GST_IS_SHARED_TASK_POOL(p) ((p) != NULL)
#define GST_SHARED_TASK_POOL(p) ((GstSharedTaskPool *) (p))
GstTaskPool *gst_shared_task_pool_new (void);
void gst_shared_task_pool_set_max_threads (GstSharedTaskPool * p, guint n);
guint gst_shared_task_pool_get_max_threads (GstSharedTaskPool * p);
void gst_task_pool_prepare (GstTaskPool * p, GError ** e);
void gst_task_pool_cleanup (GstTaskPool * p);
gpointer gst_task_pool_push (GstTaskPool * p, GstTaskPoolFunction f, gpointer d, GError ** e);
void gst_task_pool_join (GstTaskPool * p, gpointer id);

/* ---- GstVecDeque -------------------------------------------------------------------- */
typedef struct _GstVecDeque GstVecDeque;
GstVecDeque *gst_vec_deque_new (gsize initial);
void gst_vec_deque_free (GstVecDeque * d);
void gst_vec_deque_push_tail (GstVecDeque * d, gpointer p);
gpointer gst_vec_deque_pop_head (GstVecDeque * d);
gboolean gst_vec_deque_is_empty (GstVecDeque * d);

G_END_DECLS
#endif /* B200_REFSHIM_GST_H */
