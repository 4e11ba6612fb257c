#pragma once
// Middlebury imageLib stand-in (types only; nothing on the variational path uses them)
class CFloatImage {};
class CByteImage {};
