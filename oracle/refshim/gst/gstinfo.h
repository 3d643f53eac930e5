/* refshim: <gst/gstinfo.h> — logging is compiled out (see gst/gst.h) */
#include <gst/gst.h>
