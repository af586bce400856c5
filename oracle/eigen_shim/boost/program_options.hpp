// stub: Boost is not available in the build container; nothing of it is needed by the hot-path pin
