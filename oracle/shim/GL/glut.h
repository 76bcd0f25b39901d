/* Headless stand-in for <GL/glut.h>: no-op immediate-mode GL so that the reference's
 * preview/debug draw code (mesh.cpp:53-90, raytracing.cpp:422-451,518-529) links
 * without OpenGL. None of these calls are on the ray-tracing path.
 * TEST INFRASTRUCTURE ONLY (oracle/_ref build). */
#pragma once
typedef unsigned int GLenum;
typedef double GLdouble;
typedef float GLfloat;
typedef int GLint;
typedef int GLsizei;
enum { GL_POINTS = 0, GL_LINES = 1, GL_TRIANGLES = 4, GL_FRONT_AND_BACK = 0x408, GL_LINE = 0x1B01,
       GL_FILL = 0x1B02, GL_LIGHTING = 0x0B50, GL_ALL_ATTRIB_BITS = 0xFFFFF };
static inline void glBegin(GLenum) {}
static inline void glEnd() {}
static inline void glColor3f(float, float, float) {}
static inline void glColor3fv(const float*) {}
static inline void glVertex3f(float, float, float) {}
static inline void glVertex3fv(const float*) {}
static inline void glNormal3f(float, float, float) {}
static inline void glPushAttrib(unsigned) {}
static inline void glPopAttrib() {}
static inline void glDisable(GLenum) {}
static inline void glEnable(GLenum) {}
static inline void glPointSize(float) {}
static inline void glPolygonMode(GLenum, GLenum) {}
