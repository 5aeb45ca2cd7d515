// Stand-in for the reference's util/logger.hpp (spdlog singleton): the one call the drop-in shim makes,
// logger()->trace(fmt, args...), with {}-placeholders. Prints to stderr when SECEDO_B200_LOG=trace, like the
// reference with --log_level=trace. When building inside SECEDO its own header (spdlog) is used instead.
#pragma once

#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

struct StandInLogger {
    template <typename... Args>
    void trace(const std::string &fmt, Args &&...args) {
        const char *lvl = std::getenv("SECEDO_B200_LOG");
        if (!lvl || std::string(lvl) != "trace") {
            return;
        }
        std::ostringstream os;
        size_t pos = 0;
        auto emit = [&](auto &&v) {
            const size_t at = fmt.find("{}", pos);
            if (at == std::string::npos) {
                return;
            }
            os << fmt.substr(pos, at - pos) << v;
            pos = at + 2;
        };
        (void)emit;
        (emit(args), ...);
        os << fmt.substr(pos);
        std::cerr << "[trace] " << os.str() << std::endl;
    }
    StandInLogger *operator->() { return this; }
};

inline StandInLogger logger() { return StandInLogger(); }
