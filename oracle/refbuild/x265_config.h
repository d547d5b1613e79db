/* Stand-in for the header the reference's build system generates from
 * source/x265_config.h.in (one define).  Value: source/CMakeLists.txt:32. */
#ifndef X265_CONFIG_H
#define X265_CONFIG_H
#define X265_BUILD 209
#endif
