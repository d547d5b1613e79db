/* slicetype_cuda.h -- what the ENABLE_CUDA build of the reference's Lookahead (integration/x265_enable_cuda.patch on
 * source/encoder/slicetype.cpp) forwards to.  One function per public entry point of `class Lookahead`
 * (source/encoder/slicetype.h:214-227); bodies in slicetype_cuda.cpp. */
#ifndef X265_SLICETYPE_CUDA_H
#define X265_SLICETYPE_CUDA_H
#if ENABLE_CUDA

namespace X265_NS {

class Lookahead;
class Frame;

bool   cudaLookaheadCreate(Lookahead& self);                                /* Lookahead::create,  slicetype.cpp:1148 */
void   cudaLookaheadDestroy(Lookahead& self);                               /* Lookahead::destroy, :1176 */
void   cudaLookaheadAddPicture(Lookahead& self, Frame& f, int sliceType);   /* Lookahead::addPicture, :1200 */
void   cudaLookaheadFlush(Lookahead& self);                                 /* Lookahead::flush, :1246 */
Frame* cudaLookaheadGetDecided(Lookahead& self);                            /* Lookahead::getDecidedPicture, :1289 */
void   cudaLookaheadEstimatedPictureCost(Lookahead& self, Frame* cur);      /* Lookahead::getEstimatedPictureCost, :1327 */
int    cudaLookaheadFindSliceType(Lookahead& self, int poc);                /* Lookahead::findSliceType, :3248 */

}

#endif
#endif
