/* oracle/refshim/shim.c — TEST INFRASTRUCTURE ONLY.
 *
 * Implementations behind oracle/refshim/{glib.h,gst/gst.h}: the small runtime
 * the reference's arithmetic sources need once GLib / GStreamer core are
 * absent.  Written from scratch.  API the pixel/sample paths never reach
 * (caps, GValue lists, flags classes, metas) aborts loudly instead of
 * pretending to work.
 */
#include <gst/gst.h>
#include <unistd.h>

#define REFSHIM_ABORT(name) do { fprintf (stderr, "refshim: %s() is not implemented (not on the arithmetic path)\n", name); abort (); } while (0)

void
refshim_critical (const char *file, int line, const char *expr)
{
  fprintf (stderr, "refshim CRITICAL %s:%d: assertion '%s' failed\n", file, line, expr);
}

/* ------------------------------------------------------------------ strings */
gchar *
g_strdup_printf (const gchar * fmt, ...)
{
  va_list ap;
  char *out = NULL;
  va_start (ap, fmt);
  if (vasprintf (&out, fmt, ap) < 0)
    out = NULL;
  va_end (ap);
  return out;
}

GString *
g_string_new (const gchar * init)
{
  GString *s = g_new0 (GString, 1);
  s->str = strdup (init ? init : "");
  s->len = strlen (s->str);
  return s;
}

GString *
g_string_append (GString * s, const gchar * v)
{
  gsize n = strlen (v);
  s->str = realloc (s->str, s->len + n + 1);
  memcpy (s->str + s->len, v, n + 1);
  s->len += n;
  return s;
}

gchar *
g_string_free (GString * s, gboolean free_segment)
{
  gchar *r = s->str;
  if (free_segment) {
    free (r);
    r = NULL;
  }
  free (s);
  return r;
}

gchar **
g_strsplit (const gchar * s, const gchar * delim, gint max)
{
  gsize dl = strlen (delim), n = 0, cap = 8;
  gchar **v = g_new0 (gchar *, cap);
  const gchar *p = s, *q;
  (void) max;
  while ((q = strstr (p, delim)) != NULL) {
    if (n + 2 > cap)
      v = g_renew (gchar *, v, cap *= 2);
    v[n++] = strndup (p, (size_t) (q - p));
    p = q + dl;
  }
  if (n + 2 > cap)
    v = g_renew (gchar *, v, cap + 2);
  v[n++] = strdup (p);
  v[n] = NULL;
  return v;
}

void
g_strfreev (gchar ** v)
{
  gsize i;
  if (!v)
    return;
  for (i = 0; v[i]; i++)
    free (v[i]);
  free (v);
}

/* -------------------------------------------------------------- once/threads */
static pthread_mutex_t once_lock = PTHREAD_MUTEX_INITIALIZER;

gpointer
refshim_once (GOnce * once, GThreadFunc func, gpointer arg)
{
  pthread_mutex_lock (&once_lock);
  if (!once->status) {
    once->retval = func (arg);
    once->status = 1;
  }
  pthread_mutex_unlock (&once_lock);
  return once->retval;
}

guint
g_get_num_processors (void)
{
  long n = sysconf (_SC_NPROCESSORS_ONLN);
  return n > 0 ? (guint) n : 1;
}

/* ----------------------------------------------------------------- GPtrArray */
GPtrArray *
g_ptr_array_new (void)
{
  return g_new0 (GPtrArray, 1);
}

void
g_ptr_array_unref (GPtrArray * a)
{
  free (a->pdata);
  free (a);
}

static void
ptr_array_reserve (GPtrArray * a, guint n)
{
  if (n > a->alloc) {
    guint na = a->alloc ? a->alloc : 16;
    while (na < n)
      na *= 2;
    a->pdata = g_renew (gpointer, a->pdata, na);
    memset (a->pdata + a->alloc, 0, sizeof (gpointer) * (na - a->alloc));
    a->alloc = na;
  }
}

void
g_ptr_array_set_size (GPtrArray * a, gint len)
{
  ptr_array_reserve (a, (guint) len);
  if ((guint) len > a->len)
    memset (a->pdata + a->len, 0, sizeof (gpointer) * ((guint) len - a->len));
  a->len = (guint) len;
}

void
g_ptr_array_add (GPtrArray * a, gpointer p)
{
  ptr_array_reserve (a, a->len + 1);
  a->pdata[a->len++] = p;
}

void
g_ptr_array_remove_range (GPtrArray * a, guint index, guint len)
{
  g_return_if_fail (index + len <= a->len);
  memmove (a->pdata + index, a->pdata + index + len,
      sizeof (gpointer) * (a->len - index - len));
  a->len -= len;
}

/* ------------------------------------------------------------ GValue & types */
GValue *
g_value_init (GValue * v, GType t)
{
  memset (v, 0, sizeof (*v));
  v->g_type = t;
  return v;
}

void
g_value_unset (GValue * v)
{
  memset (v, 0, sizeof (*v));
}

void
g_value_set_static_string (GValue * v, const gchar * s)
{
  v->data[0].p = (gpointer) s;
}

const gchar *
g_value_get_string (const GValue * v)
{
  return (const gchar *) v->data[0].p;
}

gpointer g_type_class_ref (GType t) { (void) t; REFSHIM_ABORT ("g_type_class_ref"); return NULL; }
void g_type_class_unref (gpointer k) { (void) k; }
GFlagsValue *g_flags_get_value_by_nick (GFlagsClass * k, const gchar * n) { (void) k; (void) n; REFSHIM_ABORT ("g_flags_get_value_by_nick"); return NULL; }
GFlagsValue *g_flags_get_first_value (GFlagsClass * k, guint v) { (void) k; (void) v; REFSHIM_ABORT ("g_flags_get_first_value"); return NULL; }

/* --------------------------------------------------------------- GstStructure */
typedef struct
{
  gchar *name;
  GValue value;
  gchar *sval;                  /* owned copy for strings */
} Field;

struct _GstStructure
{
  gchar *name;
  Field *fields;
  guint n, cap;
};

GstStructure *
gst_structure_new_empty (const gchar * name)
{
  GstStructure *s = g_new0 (GstStructure, 1);
  s->name = strdup (name ? name : "");
  return s;
}

GstStructure *
gst_structure_new_static_str_empty (const gchar * name)
{
  return gst_structure_new_empty (name);
}

static Field *
find_field (const GstStructure * s, const gchar * f)
{
  guint i;
  if (!s)
    return NULL;
  for (i = 0; i < s->n; i++)
    if (strcmp (s->fields[i].name, f) == 0)
      return &s->fields[i];
  return NULL;
}

static void
set_field (GstStructure * s, const gchar * f, const GValue * v)
{
  Field *fl = find_field (s, f);
  if (!fl) {
    if (s->n == s->cap)
      s->fields = g_renew (Field, s->fields, s->cap = s->cap ? s->cap * 2 : 16);
    fl = &s->fields[s->n++];
    memset (fl, 0, sizeof (*fl));
    fl->name = strdup (f);
  }
  free (fl->sval);
  fl->sval = NULL;
  fl->value = *v;
  if (v->g_type == G_TYPE_STRING && v->data[0].p) {
    fl->sval = strdup ((const char *) v->data[0].p);
    fl->value.data[0].p = fl->sval;
  }
}

static void
set_valist (GstStructure * s, const gchar * field, va_list ap)
{
  while (field) {
    GType t = va_arg (ap, GType);
    GValue v;
    g_value_init (&v, t);
    if (t == G_TYPE_DOUBLE)
      v.data[0].d = va_arg (ap, gdouble);
    else if (t == G_TYPE_STRING)
      v.data[0].p = va_arg (ap, gpointer);
    else if (t == G_TYPE_UINT)
      v.data[0].u = va_arg (ap, guint);
    else if (t == G_TYPE_INT64 || t == G_TYPE_UINT64)
      v.data[0].i64 = va_arg (ap, gint64);
    else if (t == G_TYPE_INT || t == G_TYPE_BOOLEAN || t >= REFSHIM_TYPE_ENUM_BASE)
      v.data[0].i = va_arg (ap, gint);
    else {
      fprintf (stderr, "refshim: gst_structure_set: unsupported GType %lu for '%s'\n", t, field);
      abort ();
    }
    set_field (s, field, &v);
    field = va_arg (ap, const gchar *);
  }
}

GstStructure *
gst_structure_new (const gchar * name, const gchar * firstfield, ...)
{
  GstStructure *s = gst_structure_new_empty (name);
  va_list ap;
  va_start (ap, firstfield);
  set_valist (s, firstfield, ap);
  va_end (ap);
  return s;
}

void
gst_structure_set (GstStructure * s, const gchar * field, ...)
{
  va_list ap;
  va_start (ap, field);
  set_valist (s, field, ap);
  va_end (ap);
}

void
gst_structure_set_static_str (GstStructure * s, const gchar * field, ...)
{
  va_list ap;
  va_start (ap, field);
  set_valist (s, field, ap);
  va_end (ap);
}

GstStructure *
gst_structure_copy (const GstStructure * s)
{
  GstStructure *c = gst_structure_new_empty (s->name);
  guint i;
  for (i = 0; i < s->n; i++)
    set_field (c, s->fields[i].name, &s->fields[i].value);
  return c;
}

void
gst_structure_free (GstStructure * s)
{
  guint i;
  if (!s)
    return;
  for (i = 0; i < s->n; i++) {
    free (s->fields[i].name);
    free (s->fields[i].sval);
  }
  free (s->fields);
  free (s->name);
  free (s);
}

void
gst_structure_id_str_set_value (GstStructure * s, const GstIdStr * f, const GValue * v)
{
  set_field (s, f->s, v);
}

gboolean
gst_structure_foreach_id_str (const GstStructure * s, GstStructureForeachIdStrFunc func,
    gpointer user_data)
{
  guint i;
  for (i = 0; i < s->n; i++) {
    GstIdStr id = { s->fields[i].name };
    if (!func (&id, &s->fields[i].value, user_data))
      return FALSE;
  }
  return TRUE;
}

/* typed getters are strict about the stored GType, like the real ones */
gboolean
gst_structure_get_int (const GstStructure * s, const gchar * f, gint * v)
{
  Field *fl = find_field (s, f);
  if (!fl || fl->value.g_type != G_TYPE_INT)
    return FALSE;
  *v = fl->value.data[0].i;
  return TRUE;
}

gboolean
gst_structure_get_uint (const GstStructure * s, const gchar * f, guint * v)
{
  Field *fl = find_field (s, f);
  if (!fl || fl->value.g_type != G_TYPE_UINT)
    return FALSE;
  *v = fl->value.data[0].u;
  return TRUE;
}

gboolean
gst_structure_get_double (const GstStructure * s, const gchar * f, gdouble * v)
{
  Field *fl = find_field (s, f);
  if (!fl || fl->value.g_type != G_TYPE_DOUBLE)
    return FALSE;
  *v = fl->value.data[0].d;
  return TRUE;
}

gboolean
gst_structure_get_boolean (const GstStructure * s, const gchar * f, gboolean * v)
{
  Field *fl = find_field (s, f);
  if (!fl || fl->value.g_type != G_TYPE_BOOLEAN)
    return FALSE;
  *v = fl->value.data[0].i;
  return TRUE;
}

gboolean
gst_structure_get_enum (const GstStructure * s, const gchar * f, GType t, gint * v)
{
  Field *fl = find_field (s, f);
  if (!fl || fl->value.g_type != t)
    return FALSE;
  *v = fl->value.data[0].i;
  return TRUE;
}

const gchar *
gst_structure_get_string (const GstStructure * s, const gchar * f)
{
  Field *fl = find_field (s, f);
  if (!fl || fl->value.g_type != G_TYPE_STRING)
    return NULL;
  return (const gchar *) fl->value.data[0].p;
}

const gchar *gst_structure_get_name (const GstStructure * s) { return s->name; }
gboolean gst_structure_has_name (const GstStructure * s, const gchar * n) { return strcmp (s->name, n) == 0; }
gboolean gst_structure_has_field (const GstStructure * s, const gchar * f) { return find_field (s, f) != NULL; }
gboolean gst_structure_get_fraction (const GstStructure * s, const gchar * f, gint * n, gint * d) { (void) s; (void) f; (void) n; (void) d; return FALSE; }
gboolean gst_structure_get_flagset (const GstStructure * s, const gchar * f, guint * fl, guint * m) { (void) s; (void) f; (void) fl; (void) m; return FALSE; }

/* --------------------------------------------------------- scalar core helpers */
const gchar *
gst_format_get_name (GstFormat f)
{
  static const char *n[] = { "undefined", "default", "bytes", "time", "buffers", "percent" };
  return ((guint) f < 6) ? n[f] : "?";
}

guint64
gst_util_uint64_scale (guint64 val, guint64 num, guint64 denom)
{
  return (guint64) (((unsigned __int128) val * num) / denom);
}

guint64
gst_util_uint64_scale_int (guint64 val, gint num, gint denom)
{
  return (guint64) (((unsigned __int128) val * (guint64) num) / (guint64) denom);
}

gint
gst_util_greatest_common_divisor (gint a, gint b)
{
  while (b != 0) {
    gint t = a;
    a = b;
    b = t % b;
  }
  return ABS (a);
}

/* ------------------------------------------------------------------ task pool */
typedef struct
{
  pthread_t th;
  GstTaskPoolFunction f;
  gpointer d;
} PoolTask;

static void *
pool_tramp (void *p)
{
  PoolTask *t = p;
  t->f (t->d);
  return NULL;
}

GstTaskPool *gst_shared_task_pool_new (void) { return g_new0 (GstTaskPool, 1); }
void gst_shared_task_pool_set_max_threads (GstSharedTaskPool * p, guint n) { p->max_threads = n; }
guint gst_shared_task_pool_get_max_threads (GstSharedTaskPool * p) { return p->max_threads; }
void gst_task_pool_prepare (GstTaskPool * p, GError ** e) { (void) p; (void) e; }
void gst_task_pool_cleanup (GstTaskPool * p) { (void) p; }

gpointer
gst_task_pool_push (GstTaskPool * p, GstTaskPoolFunction f, gpointer d, GError ** e)
{
  PoolTask *t = g_new0 (PoolTask, 1);
  (void) p;
  (void) e;
  t->f = f;
  t->d = d;
  if (pthread_create (&t->th, NULL, pool_tramp, t) != 0) {
    free (t);
    return NULL;
  }
  return t;
}

void
gst_task_pool_join (GstTaskPool * p, gpointer id)
{
  PoolTask *t = id;
  (void) p;
  pthread_join (t->th, NULL);
  free (t);
}

/* ---------------------------------------------------------------- GstVecDeque */
struct _GstVecDeque
{
  gpointer *d;
  gsize head, len, cap;
};

GstVecDeque *
gst_vec_deque_new (gsize initial)
{
  GstVecDeque *q = g_new0 (GstVecDeque, 1);
  q->cap = initial ? initial : 8;
  q->d = g_new0 (gpointer, q->cap);
  return q;
}

void
gst_vec_deque_free (GstVecDeque * q)
{
  free (q->d);
  free (q);
}

void
gst_vec_deque_push_tail (GstVecDeque * q, gpointer p)
{
  if (q->len == q->cap) {
    gsize i, nc = q->cap * 2;
    gpointer *nd = g_new0 (gpointer, nc);
    for (i = 0; i < q->len; i++)
      nd[i] = q->d[(q->head + i) % q->cap];
    free (q->d);
    q->d = nd;
    q->head = 0;
    q->cap = nc;
  }
  q->d[(q->head + q->len++) % q->cap] = p;
}

gpointer
gst_vec_deque_pop_head (GstVecDeque * q)
{
  gpointer p;
  if (!q->len)
    return NULL;
  p = q->d[q->head];
  q->head = (q->head + 1) % q->cap;
  q->len--;
  return p;
}

gboolean gst_vec_deque_is_empty (GstVecDeque * q) { return q->len == 0; }

/* -------------------------------------------- never-reached API: abort loudly */
gboolean gst_buffer_is_writable (GstBuffer * b) { (void) b; REFSHIM_ABORT ("gst_buffer_is_writable"); return FALSE; }
GstMeta *gst_buffer_iterate_meta (GstBuffer * b, gpointer * s) { (void) b; (void) s; REFSHIM_ABORT ("gst_buffer_iterate_meta"); return NULL; }
gboolean gst_meta_api_type_tags_contain_only (GType a, const gchar ** t) { (void) a; (void) t; REFSHIM_ABORT ("gst_meta_api_type_tags_contain_only"); return FALSE; }
gboolean gst_value_deserialize (GValue * d, const gchar * s) { (void) d; (void) s; REFSHIM_ABORT ("gst_value_deserialize"); return FALSE; }
guint gst_value_list_get_size (const GValue * v) { (void) v; REFSHIM_ABORT ("gst_value_list_get_size"); return 0; }
const GValue *gst_value_list_get_value (const GValue * v, guint i) { (void) v; (void) i; REFSHIM_ABORT ("gst_value_list_get_value"); return NULL; }
void gst_value_list_append_and_take_value (GValue * v, GValue * a) { (void) v; (void) a; REFSHIM_ABORT ("gst_value_list_append_and_take_value"); }
gboolean gst_caps_is_fixed (const GstCaps * c) { (void) c; REFSHIM_ABORT ("gst_caps_is_fixed"); return FALSE; }
GstStructure *gst_caps_get_structure (const GstCaps * c, guint i) { (void) c; (void) i; REFSHIM_ABORT ("gst_caps_get_structure"); return NULL; }
GstCapsFeatures *gst_caps_get_features (const GstCaps * c, guint i) { (void) c; (void) i; REFSHIM_ABORT ("gst_caps_get_features"); return NULL; }
gboolean gst_caps_features_contains (const GstCapsFeatures * f, const gchar * s) { (void) f; (void) s; REFSHIM_ABORT ("gst_caps_features_contains"); return FALSE; }
GstCaps *gst_caps_new_static_str_simple (const gchar * m, const gchar * f, ...) { (void) m; (void) f; REFSHIM_ABORT ("gst_caps_new_static_str_simple"); return NULL; }
void gst_caps_set_simple_static_str (GstCaps * c, const gchar * f, ...) { (void) c; (void) f; REFSHIM_ABORT ("gst_caps_set_simple_static_str"); }
void gst_caps_set_simple (GstCaps * c, const gchar * f, ...) { (void) c; (void) f; REFSHIM_ABORT ("gst_caps_set_simple"); }
GstCapsFeatures *gst_caps_features_new_static_str (const gchar * f, ...) { (void) f; REFSHIM_ABORT ("gst_caps_features_new_static_str"); return NULL; }
void gst_caps_set_features (GstCaps * c, guint i, GstCapsFeatures * f) { (void) c; (void) i; (void) f; REFSHIM_ABORT ("gst_caps_set_features"); }
GstCaps *gst_caps_new_full (gpointer s, ...) { (void) s; REFSHIM_ABORT ("gst_caps_new_full"); return NULL; }
GstStructure *gst_structure_new_static_str (const gchar * n, const gchar * f, ...) { (void) n; (void) f; REFSHIM_ABORT ("gst_structure_new_static_str"); return NULL; }
void gst_structure_take_value_static_str (GstStructure * s, const gchar * f, GValue * v) { (void) s; (void) f; (void) v; REFSHIM_ABORT ("gst_structure_take_value_static_str"); }

/* libgstvideo symbols from files we do not compile (gstvideometa.c, video-multiview.c);
 * referenced only by caps/meta helpers that the pixel path never runs. */
GQuark gst_video_meta_transform_matrix_get_quark (void) { REFSHIM_ABORT ("gst_video_meta_transform_matrix_get_quark"); return 0; }
GQuark gst_video_meta_transform_scale_get_quark (void) { REFSHIM_ABORT ("gst_video_meta_transform_scale_get_quark"); return 0; }
void gst_video_meta_transform_matrix_init (gpointer t, gconstpointer a, gconstpointer b, gconstpointer c, gconstpointer d) { (void) t; (void) a; (void) b; (void) c; (void) d; REFSHIM_ABORT ("gst_video_meta_transform_matrix_init"); }
GType gst_video_multiview_flagset_get_type (void) { return (GType) 0; }
gint gst_video_multiview_mode_from_caps_string (const gchar * s) { (void) s; REFSHIM_ABORT ("gst_video_multiview_mode_from_caps_string"); return 0; }
const gchar *gst_video_multiview_mode_to_caps_string (gint m) { (void) m; REFSHIM_ABORT ("gst_video_multiview_mode_to_caps_string"); return NULL; }
