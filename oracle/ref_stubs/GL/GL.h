/* Empty stand-in for <GL/GL.h> (Win32 spelling, included by the reference's InitShader.h).
 * GL types come from the reference's own vendored include/GL/glew.h. */
