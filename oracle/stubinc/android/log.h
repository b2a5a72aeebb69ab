// stub of the Android NDK logging header (the reference logs through it; no-op here).
#ifndef MINICV_ANDROID_LOG_H
#define MINICV_ANDROID_LOG_H
enum { ANDROID_LOG_UNKNOWN = 0, ANDROID_LOG_DEFAULT, ANDROID_LOG_VERBOSE, ANDROID_LOG_DEBUG, ANDROID_LOG_INFO,
       ANDROID_LOG_WARN, ANDROID_LOG_ERROR, ANDROID_LOG_FATAL, ANDROID_LOG_SILENT };
static inline int __android_log_print(int, const char*, const char*, ...) { return 0; }
#endif
