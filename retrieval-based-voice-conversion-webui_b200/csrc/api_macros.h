#pragma once
#define RVCB_API_BEGIN try {
#define RVCB_API_END                                   \
    }                                                  \
    catch (const std::exception& e) {                  \
        rvcb::set_last_error(e.what());                \
        return -1;                                     \
    }                                                  \
    catch (...) {                                      \
        rvcb::set_last_error("unknown C++ exception"); \
        return -2;                                     \
    }                                                  \
    return 0;
