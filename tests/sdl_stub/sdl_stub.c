/* sdl_stub.c -- a headless stand-in for the eleven SDL2 functions the reference's interactive host calls (TEST
 * INFRASTRUCTURE).  It exists so that demo-interactive/liblys.c can be LINKED AND RUN UNMODIFIED against libtracer: the
 * reference checkout ships SDL2's headers (deps/SDL2/include) but not its library blob (deps/SDL2/lib/libSDL2.a is missing)
 * and neither the build container nor the GPU box has a display.  Compiled against the reference's own SDL2 headers, so the
 * event and surface structs have SDL's layout.
 *
 * The "window" is a 32-bit pixel buffer.  Events come from a script in the environment, one event per frame boundary:
 *     LYS_SDL_SCRIPT="0:resize:64x48 0:key:109 2:keyup:109 5:quit"      (frame:kind[:argument]; key codes are SDL keycodes)
 * A frame ends at SDL_UpdateWindowSurface.  At SDL_Quit the last window contents are written to $LYS_SDL_DUMP as a binary PPM
 * (masks 0xFF0000 / 0xFF00 / 0xFF, the ones liblys.c:59 passes to SDL_CreateRGBSurfaceFrom) and a one-line summary is printed.
 */
#include <SDL2/SDL.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct SDL_Window { int w, h; SDL_Surface *surface; };

static int g_frame = 0, g_cursor = 0, g_events = 0;
static struct SDL_Window *g_wnd = NULL;

static SDL_Surface *make_surface(void *pixels, int w, int h, int pitch, int owns) {
    SDL_Surface *s = calloc(1, sizeof *s);
    s->w = w; s->h = h; s->pitch = pitch;
    s->pixels = owns ? calloc((size_t)(h > 0 ? h : 1), (size_t)pitch) : pixels;
    s->flags = owns ? 0 : SDL_PREALLOC;
    return s;
}

int SDL_Init(Uint32 flags) { (void)flags; return 0; }
const char *SDL_GetError(void) { return ""; }
SDL_Window *SDL_CreateWindow(const char *title, int x, int y, int w, int h, Uint32 flags) {
    (void)title; (void)x; (void)y; (void)flags;
    struct SDL_Window *wnd = calloc(1, sizeof *wnd);
    wnd->w = w; wnd->h = h;
    g_wnd = wnd;
    return wnd;
}
SDL_Surface *SDL_GetWindowSurface(SDL_Window *wnd) {
    if (!wnd->surface || wnd->surface->w != wnd->w || wnd->surface->h != wnd->h) {       /* a resize invalidates the old surface */
        if (wnd->surface) { free(wnd->surface->pixels); free(wnd->surface); }
        wnd->surface = make_surface(NULL, wnd->w, wnd->h, wnd->w * 4, 1);
    }
    return wnd->surface;
}
SDL_Surface *SDL_CreateRGBSurfaceFrom(void *pixels, int w, int h, int depth, int pitch, Uint32 r, Uint32 g, Uint32 b, Uint32 a) {
    (void)r; (void)g; (void)b; (void)a;
    if (depth != 32) return NULL;
    return make_surface(pixels, w, h, pitch, 0);
}
void SDL_FreeSurface(SDL_Surface *s) { if (s) { if (!(s->flags & SDL_PREALLOC)) free(s->pixels); free(s); } }
int SDL_UpperBlit(SDL_Surface *src, const SDL_Rect *sr, SDL_Surface *dst, SDL_Rect *dr) {
    (void)sr; (void)dr;
    int h = src->h < dst->h ? src->h : dst->h, w = src->w < dst->w ? src->w : dst->w;
    for (int y = 0; y < h; y++) memcpy((char *)dst->pixels + (size_t)y * dst->pitch, (char *)src->pixels + (size_t)y * src->pitch, (size_t)w * 4);
    return 0;
}
int SDL_UpdateWindowSurface(SDL_Window *wnd) { (void)wnd; g_frame++; return 0; }

/* next scripted event whose frame number has been reached */
int SDL_PollEvent(SDL_Event *ev) {
    const char *script = getenv("LYS_SDL_SCRIPT");
    if (!script) script = "3:quit";
    const char *p = script;
    for (int k = 0; ; k++) {
        while (*p == ' ') p++;
        if (!*p) return 0;
        int frame = 0, n = 0;
        char kind[16] = "", arg[32] = "";
        if (sscanf(p, "%d:%15[a-z]%n", &frame, kind, &n) < 2) { fprintf(stderr, "sdl_stub: bad script at '%s'\n", p); exit(EXIT_FAILURE); }
        p += n;
        if (*p == ':') { p++; n = 0; sscanf(p, "%31[^ ]%n", arg, &n); p += n; }
        if (k < g_cursor) continue;                       /* already delivered */
        if (frame > g_frame) return 0;                    /* not yet */
        g_cursor = k + 1; g_events++;
        memset(ev, 0, sizeof *ev);
        if (!strcmp(kind, "quit")) { ev->type = SDL_QUIT; }
        else if (!strcmp(kind, "key") || !strcmp(kind, "keyup")) {
            ev->type = !strcmp(kind, "key") ? SDL_KEYDOWN : SDL_KEYUP;
            ev->key.type = ev->type;
            ev->key.keysym.sym = (SDL_Keycode)strtol(arg, NULL, 0);
        } else if (!strcmp(kind, "resize")) {
            int w = 0, h = 0;
            if (sscanf(arg, "%dx%d", &w, &h) != 2) { fprintf(stderr, "sdl_stub: bad resize '%s'\n", arg); exit(EXIT_FAILURE); }
            g_wnd->w = w; g_wnd->h = h;
            ev->type = SDL_WINDOWEVENT;
            ev->window.event = SDL_WINDOWEVENT_RESIZED;
            ev->window.data1 = w; ev->window.data2 = h;
        } else { fprintf(stderr, "sdl_stub: unknown event kind '%s'\n", kind); exit(EXIT_FAILURE); }
        return 1;
    }
}
void SDL_DestroyWindow(SDL_Window *wnd) { (void)wnd; }      /* kept until SDL_Quit for the dump */
void SDL_Quit(void) {
    if (!g_wnd || !g_wnd->surface) return;
    SDL_Surface *s = g_wnd->surface;
    unsigned long long sum = 0;
    for (int y = 0; y < s->h; y++) for (int x = 0; x < s->w; x++) sum += ((Uint32 *)((char *)s->pixels + (size_t)y * s->pitch))[x] & 0xFFFFFFu;
    const char *dump = getenv("LYS_SDL_DUMP");
    if (dump) {
        FILE *fp = fopen(dump, "wb");
        if (fp) {
            fprintf(fp, "P6\n%d %d\n255\n", s->w, s->h);
            for (int y = 0; y < s->h; y++) for (int x = 0; x < s->w; x++) {
                Uint32 p = ((Uint32 *)((char *)s->pixels + (size_t)y * s->pitch))[x];
                unsigned char rgb[3] = {(unsigned char)(p >> 16), (unsigned char)(p >> 8), (unsigned char)p};
                fwrite(rgb, 1, 3, fp);
            }
            fclose(fp);
        }
    }
    printf("sdl_stub: %d frames, %d events, window %dx%d, rgb checksum %llu\n", g_frame, g_events, s->w, s->h, sum);
}
