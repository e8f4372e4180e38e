/* Stand-in for <GL/gl.h> on build machines without OpenGL headers (this image has none): just the two scalar types
 * cuda_gl_interop.h needs.  build.py puts this directory on the include path ONLY when no system GL/gl.h exists. */
#ifndef CWA_GL_COMPAT_H
#define CWA_GL_COMPAT_H
typedef unsigned int GLenum;
typedef unsigned int GLuint;
#ifndef GL_TEXTURE_2D
#define GL_TEXTURE_2D 0x0DE1
#endif
#endif
