// g19/compat.h -- picks the vector-math and image back ends for the interface
// headers: the reference's own dependencies when they are installed (vendored
// GLM, Qt5Gui), minimal stand-ins otherwise.
#pragma once
#if defined(__has_include)
#  if __has_include(<glm/glm.hpp>)
#    include <glm/glm.hpp>
#    define G19_HAVE_GLM 1
#  endif
#  if __has_include(<QImage>) && !defined(G19_NO_QT)
#    include <QImage>
#    include <QColor>
#    define G19_HAVE_QT 1
#  endif
#endif
#ifndef G19_HAVE_GLM
#  include "g19/vecmath_min.h"
#endif
#ifndef G19_HAVE_QT
#  include "g19/qimage_min.h"
#endif
#include "g19.h"
