/*
 * Headless SDL-1.2 stand-in (implementation). TEST INFRASTRUCTURE ONLY -- see SDL/SDL.h.
 *
 * Behaviour relevant to the oracle:
 *   - SDL_Quit() dumps the reference's global framebuffer `vfb` (src/main.cpp:53) as
 *     int32 w, int32 h, float32 rgb[h][w][3] to the file named by $FRAY_DUMP.
 *   - SDL_Delay() sleeps at most 1 ms so the 100 ms poll in renderScene_threaded()
 *     (src/sdl.cpp:224) does not quantise the reference's own "Render took" figure.
 *   - SDL_ThreadID() returns a small per-thread ordinal (the reference keys its RNG table by it,
 *     src/random_generator.cpp:128-131).
 */
#include "SDL/SDL.h"
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <atomic>

struct SDL_Thread { pthread_t t; int (*fn)(void*); void* data; int status; };
struct SDL_mutex { pthread_mutex_t m; };
struct SDL_cond { pthread_cond_t c; };

extern float vfb[]; // Color vfb[VFB_MAX_SIZE][VFB_MAX_SIZE], src/main.cpp:53 (3 floats per pixel)
#ifndef FRAY_REF_VFB_STRIDE
#define FRAY_REF_VFB_STRIDE 3000 // VFB_MAX_SIZE, src/constants.h:27
#endif

static SDL_Surface g_screen;
static SDL_PixelFormat g_format;
static Uint8 g_keys[SDLK_LAST];
static struct timespec g_t0;
static char g_err[] = "headless shim";

extern "C" {

int SDL_Init(Uint32) { clock_gettime(CLOCK_MONOTONIC, &g_t0); return 0; }

void SDL_Quit(void)
{
	const char* fn = getenv("FRAY_DUMP");
	if (!fn || !g_screen.pixels) return;
	FILE* f = fopen(fn, "wb");
	if (!f) { fprintf(stderr, "shim: cannot write %s\n", fn); return; }
	int32_t wh[2] = { g_screen.w, g_screen.h };
	fwrite(wh, sizeof(wh), 1, f);
	for (int y = 0; y < g_screen.h; y++)
		fwrite(vfb + (size_t) y * FRAY_REF_VFB_STRIDE * 3, sizeof(float) * 3, g_screen.w, f);
	fclose(f);
}

char* SDL_GetError(void) { return g_err; }

SDL_Surface* SDL_SetVideoMode(int w, int h, int, Uint32)
{
	g_format.BitsPerPixel = 32; g_format.BytesPerPixel = 4;
	g_format.Rshift = 16; g_format.Gshift = 8; g_format.Bshift = 0; g_format.Ashift = 24;
	g_screen.format = &g_format;
	g_screen.w = w; g_screen.h = h;
	g_screen.pitch = (Uint16) (w * 4);
	g_screen.pixels = calloc((size_t) w * h, 4);
	return &g_screen;
}

int SDL_Flip(SDL_Surface*) { return 0; }
void SDL_UpdateRect(SDL_Surface*, int, int, Uint32, Uint32) {}
void SDL_WM_SetCaption(const char*, const char*) {}
int SDL_ShowCursor(int) { return 0; }

Uint32 SDL_GetTicks(void)
{
	struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t);
	return (Uint32) ((t.tv_sec - g_t0.tv_sec) * 1000 + (t.tv_nsec - g_t0.tv_nsec) / 1000000);
}

void SDL_Delay(Uint32 ms) { usleep(ms ? 1000 : 0); }
int SDL_PollEvent(SDL_Event*) { return 0; }
int SDL_WaitEvent(SDL_Event* ev) { memset(ev, 0, sizeof(*ev)); ev->type = SDL_QUIT; return 1; }
Uint8* SDL_GetKeyState(int* n) { if (n) *n = SDLK_LAST; return g_keys; }
Uint8 SDL_GetRelativeMouseState(int* x, int* y) { if (x) *x = 0; if (y) *y = 0; return 0; }

static void* thread_tramp(void* p)
{
	SDL_Thread* t = (SDL_Thread*) p;
	t->status = t->fn(t->data);
	return NULL;
}

SDL_Thread* SDL_CreateThread(int (*fn)(void*), void* data)
{
	SDL_Thread* t = (SDL_Thread*) calloc(1, sizeof(SDL_Thread));
	t->fn = fn; t->data = data;
	if (pthread_create(&t->t, NULL, thread_tramp, t)) { free(t); return NULL; }
	return t;
}

void SDL_WaitThread(SDL_Thread* t, int* status)
{
	if (!t) return;
	pthread_join(t->t, NULL);
	if (status) *status = t->status;
	free(t);
}

Uint32 SDL_ThreadID(void)
{
	static std::atomic<Uint32> next(1);
	static thread_local Uint32 mine = 0;
	if (!mine) mine = next++;
	return mine;
}

SDL_mutex* SDL_CreateMutex(void)
{
	SDL_mutex* m = (SDL_mutex*) malloc(sizeof(SDL_mutex));
	pthread_mutexattr_t a;
	pthread_mutexattr_init(&a);
	pthread_mutexattr_settype(&a, PTHREAD_MUTEX_RECURSIVE); // SDL-1.2 mutexes are recursive
	pthread_mutex_init(&m->m, &a);
	pthread_mutexattr_destroy(&a);
	return m;
}
void SDL_DestroyMutex(SDL_mutex* m) { if (m) { pthread_mutex_destroy(&m->m); free(m); } }
int SDL_mutexP(SDL_mutex* m) { return m ? pthread_mutex_lock(&m->m) : -1; }
int SDL_mutexV(SDL_mutex* m) { return m ? pthread_mutex_unlock(&m->m) : -1; }

SDL_cond* SDL_CreateCond(void)
{
	SDL_cond* c = (SDL_cond*) malloc(sizeof(SDL_cond));
	pthread_cond_init(&c->c, NULL);
	return c;
}
void SDL_DestroyCond(SDL_cond* c) { if (c) { pthread_cond_destroy(&c->c); free(c); } }
int SDL_CondWait(SDL_cond* c, SDL_mutex* m) { return pthread_cond_wait(&c->c, &m->m); }
int SDL_CondSignal(SDL_cond* c) { return pthread_cond_signal(&c->c); }
int SDL_CondBroadcast(SDL_cond* c) { return pthread_cond_broadcast(&c->c); }

} // extern "C"
