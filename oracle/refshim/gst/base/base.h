/* refshim: <gst/base/base.h> — nothing from libgstbase is used on the arithmetic path. */
#include <gst/gst.h>
