// Drop-in for include/efanna2e/exceptions.h.
#pragma once
#include <stdexcept>

namespace efanna2e {
class NotImplementedException : public std::logic_error {
   public:
    NotImplementedException() : std::logic_error("Function not yet implemented.") {}
};
}  // namespace efanna2e
