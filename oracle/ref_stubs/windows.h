/* Empty stand-in for <windows.h>: the reference's InitShader.h includes it (Win32 app); the
 * CPU grid twin we compile from the reference sources uses nothing from it. */
