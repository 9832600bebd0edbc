"""Second, independent restatement of raisin's LZSS + Huffman path: a statement-by-statement
Python transcription of the Go code, kept deliberately naive (same control flow, same data
structures, same loops) so that it can cross-check oracle/raisin_oracle.c on small inputs.

TEST INFRASTRUCTURE ONLY.  Pure-Python loops: use on inputs of a few KiB.

Transcribed from (paths relative to the reference repository root):
  compressor/lz/lzss.go          109-184, 224-320, 323-364, 366-406, 418-433
  compressor/huffman/huffman.go  35-54, 58-103, 110-127, 131-153, 174-191, 196-227, 229-256,
                                 258-297, 299-325
Go stdlib pieces restated: bytes.Index, container/heap (Init/Push/Pop/up/down), strconv.Atoi,
range-over-string rune decoding, string(rune).
"""
from __future__ import annotations

# ----------------------------------------------------------------------------- Go stdlib


class GoPanic(Exception):
    """A Go runtime panic (index out of range, explicit panic, ...)."""


def go_atoi(s: bytes) -> int:
    """strconv.Atoi with the error dropped (callers use `v, _ :=`)."""
    if len(s) == 0:
        return 0
    neg = False
    t = s
    if t[0:1] in (b"+", b"-"):
        neg = t[0:1] == b"-"
        t = t[1:]
        if len(t) == 0:
            return 0
    maxv = (1 << 64) - 1
    cutoff = maxv // 10 + 1
    un = 0
    for ch in t:
        if ch < 0x30 or ch > 0x39:
            return 0
        if un >= cutoff:
            un = maxv
            break
        un *= 10
        n1 = un + (ch - 0x30)
        if n1 > maxv:
            un = maxv
            break
        un = n1
    if not neg and un >= (1 << 63):
        return (1 << 63) - 1
    if neg and un > (1 << 63):
        return -(1 << 63)
    return -un if neg else un


def go_range_string(b: bytes):
    """`for i, c := range string(b)`: yields (byte index, rune)."""
    i, n = 0, len(b)
    while i < n:
        c = b[i]
        if c < 0x80:
            yield i, c
            i += 1
            continue
        lo, hi = 0x80, 0xBF
        if 0xC2 <= c <= 0xDF:
            need = 2
        elif c == 0xE0:
            need, lo = 3, 0xA0
        elif 0xE1 <= c <= 0xEC or c in (0xEE, 0xEF):
            need = 3
        elif c == 0xED:
            need, hi = 3, 0x9F
        elif c == 0xF0:
            need, lo = 4, 0x90
        elif 0xF1 <= c <= 0xF3:
            need = 4
        elif c == 0xF4:
            need, hi = 4, 0x8F
        else:
            yield i, 0xFFFD
            i += 1
            continue
        ok = i + need <= n and lo <= b[i + 1] <= hi
        if ok:
            for k in range(2, need):
                if not (0x80 <= b[i + k] <= 0xBF):
                    ok = False
                    break
        if not ok:
            yield i, 0xFFFD
            i += 1
            continue
        if need == 2:
            r = ((c & 0x1F) << 6) | (b[i + 1] & 0x3F)
        elif need == 3:
            r = ((c & 0x0F) << 12) | ((b[i + 1] & 0x3F) << 6) | (b[i + 2] & 0x3F)
        else:
            r = ((c & 0x07) << 18) | ((b[i + 1] & 0x3F) << 12) | ((b[i + 2] & 0x3F) << 6) | (b[i + 3] & 0x3F)
        yield i, r
        i += need


def go_string_rune(r: int) -> bytes:
    """string(rune)."""
    if r < 0 or r > 0x10FFFF or 0xD800 <= r <= 0xDFFF:
        r = 0xFFFD
    return chr(r).encode("utf-8")


# ----------------------------------------------------------------------------- lzss.go


def EncodeOpeningSymbols(data: bytes) -> bytes:  # lzss.go:369-389
    encoded = bytearray()
    foundEscape = False
    for val in data:
        if val == 0x3C:
            if foundEscape:
                encoded.append(0x5C)
            val = 0xFF
        elif val == 0xFF or val == 0x5C:
            encoded.append(0x5C)
        elif val == 0x5C:  # unreachable, as in the reference
            if foundEscape:
                encoded.append(0x5C)
            foundEscape = True
        encoded.append(val)
    return bytes(encoded)


def DecodeOpeningSymbols(data: bytes) -> bytes:  # lzss.go:391-406
    decoded = bytearray()
    foundEscape = False
    for val in data:
        if val == 0xFF and not foundEscape:
            decoded += b"<"
        elif val == 0x5C and not foundEscape:
            foundEscape = True
        else:
            foundEscape = False
            decoded.append(val)
    return bytes(decoded)


def getEncoding(relativePointer: int, relativeOffset: int) -> bytes:  # lzss.go:318-320
    return b"<" + str(relativePointer).encode() + b"," + str(relativeOffset).encode() + b">"


def FindReverseSlice(inp: bytes, val: bytes):  # lzss.go:418-421
    index = inp.find(val)
    return index, index != -1


def FindReverse(sl: bytes, val: int):  # lzss.go:423-433
    i = len(sl) - 1
    while i >= 0:
        if sl[i] == val:
            return i, True
        i -= 1  # the extra decrement inside the loop body
        i -= 1  # the loop's own i--
    return -1, False


class Reference:
    __slots__ = ("value", "isReference", "negativeOffset", "size")

    def __init__(self, value, isReference=False, negativeOffset=0, size=0):
        self.value, self.isReference, self.negativeOffset, self.size = value, isReference, negativeOffset, size


def compressorWorker(searchBuffer: bytes, scanBytes: bytes, nextBytes: bytes) -> Reference:  # lzss.go:166-184
    # recursion unrolled into a loop over growing scanBytes; returns what the outermost call returns
    best = None
    while True:
        index, found = FindReverseSlice(searchBuffer, scanBytes)
        if not found:
            # Reference{value: scanBytes}: the caller keeps its own level (or this is level 1)
            return best if best is not None else Reference(scanBytes)
        best = Reference(scanBytes, True, len(searchBuffer) - index, len(scanBytes))
        if len(nextBytes) > 0:
            scanBytes = scanBytes + nextBytes[0:1]
            nextBytes = nextBytes[1:]
        else:
            return best


def CompressAsync(fileContents: bytes, maxSearchBufferLength: int) -> bytes:  # lzss.go:109-154
    fileContents = EncodeOpeningSymbols(fileContents)
    output = []
    for i in range(len(fileContents)):
        startIndex = 0
        searchBuffer = fileContents[:i]
        if maxSearchBufferLength > 0 and len(searchBuffer) > maxSearchBufferLength:
            startIndex = len(searchBuffer) - maxSearchBufferLength
        # compressorWorkerAsync passes nextBytes[1:] (lzss.go:159)
        output.append(compressorWorker(searchBuffer[startIndex:], fileContents[i : i + 1], fileContents[i:][1:]))
    finalOutput = bytearray()
    ignoreNextChars = 0
    for ref in output:
        if ignoreNextChars > 0:
            ignoreNextChars -= 1
        elif ref.isReference:
            ignoreNextChars = ref.size - 1
            if len(getEncoding(ref.negativeOffset, ref.size)) < ref.size:
                finalOutput += getEncoding(ref.negativeOffset, ref.size)
            else:
                finalOutput += ref.value
        else:
            finalOutput += ref.value
    return bytes(finalOutput)


def Compress(fileContents: bytes, maxSearchBufferLength: int) -> bytes:  # lzss.go:224-316
    fileContents = EncodeOpeningSymbols(fileContents)
    searchBuffer = bytearray()
    output = bytearray()
    pointer = 0
    checkNextByte = False
    checkStartPointer = 0
    checkOffset = 0
    checkBytesToAdd = bytearray()
    MinimumSizeOfReference = -1
    for fileByte in fileContents:
        index, found = 0, False
        if not checkNextByte:
            index, found = FindReverse(searchBuffer, fileByte)
        else:
            diminishingReturns = 0
            if maxSearchBufferLength > 0 and len(searchBuffer) > maxSearchBufferLength:
                diminishingReturns = len(searchBuffer) - maxSearchBufferLength
            index, found = FindReverseSlice(
                bytes(searchBuffer[diminishingReturns:]), bytes(checkBytesToAdd) + bytes([fileByte])
            )
        if found and checkNextByte:
            pointer = len(searchBuffer) - index
            checkStartPointer = pointer
            checkOffset += 1
            checkBytesToAdd.append(fileByte)
        elif found and not checkNextByte:
            pointer = len(searchBuffer) - index
            checkStartPointer = pointer
            checkOffset = 1
            checkNextByte = True
            checkBytesToAdd.append(fileByte)
        else:
            if checkNextByte:
                shouldAdd = True
                if MinimumSizeOfReference == -1:
                    if len(getEncoding(checkStartPointer, checkOffset)) > len(checkBytesToAdd):
                        shouldAdd = False
                if len(checkBytesToAdd) > MinimumSizeOfReference and shouldAdd:
                    output += getEncoding(checkStartPointer, checkOffset)
                else:
                    output += checkBytesToAdd
                checkStartPointer = 0
                checkOffset = 0
                checkNextByte = False
                searchBuffer += checkBytesToAdd
                checkBytesToAdd = bytearray()
            output.append(fileByte)
        if not checkNextByte:
            searchBuffer.append(fileByte)
    if checkNextByte:
        shouldAdd = True
        if MinimumSizeOfReference == -1:
            if len(getEncoding(checkStartPointer, checkOffset)) > len(checkBytesToAdd):
                shouldAdd = False
        if len(checkBytesToAdd) > MinimumSizeOfReference and shouldAdd:
            output += getEncoding(checkStartPointer, checkOffset)
        else:
            output += checkBytesToAdd
    return bytes(output)


def Decompress(fileContents: bytes) -> bytes:  # lzss.go:323-364
    searchBuffer = bytearray()
    output = bytearray()
    pointer = 0
    pointerBytes = bytearray()
    offset = 0
    offsetBytes = bytearray()
    lookingFor = "<"
    for fileByte in fileContents:
        if lookingFor == "<" and fileByte == 0x3C:
            lookingFor = ","
        elif lookingFor == ",":
            if fileByte == 0x2C:
                lookingFor = ">"
                pointer = go_atoi(bytes(pointerBytes))
                pointerBytes = bytearray()
            else:
                pointerBytes.append(fileByte)
        elif lookingFor == ">":
            if fileByte == 0x3E:
                lookingFor = "<"
                offset = go_atoi(bytes(offsetBytes))
                offsetBytes = bytearray()
                absolutePointer = len(searchBuffer) - pointer
                lo, hi = absolutePointer, absolutePointer + offset
                if lo < 0 or hi < lo or hi > len(searchBuffer):  # Go: slice bounds panic / slack read
                    raise GoPanic("slice bounds out of range")
                sl = bytes(searchBuffer[lo:hi])
                output += sl
                searchBuffer += sl
            else:
                offsetBytes.append(fileByte)
        else:
            output.append(fileByte)
            searchBuffer.append(fileByte)
    return DecodeOpeningSymbols(bytes(output))


# ----------------------------------------------------------------------------- huffman.go


class HuffmanLeaf:
    __slots__ = ("freq", "value")

    def __init__(self, freq, value):
        self.freq, self.value = freq, value

    def Freq(self):
        return self.freq


class HuffmanNode:
    __slots__ = ("freq", "left", "right")

    def __init__(self, freq, left, right):
        self.freq, self.left, self.right = freq, left, right

    def Freq(self):
        return self.freq


class treeHeap(list):  # huffman.go:40-54
    def Len(self):
        return len(self)

    def Less(self, i, j):
        return self[i].Freq() < self[j].Freq()

    def Swap(self, i, j):
        self[i], self[j] = self[j], self[i]

    def Push(self, e):
        self.append(e)

    def Pop(self):
        return self.pop()


def _heap_up(h, j):  # container/heap.up
    while True:
        i = int((j - 1) / 2)  # Go truncating division
        if i == j or not h.Less(j, i):
            break
        h.Swap(i, j)
        j = i


def _heap_down(h, i0, n):  # container/heap.down
    i = i0
    while True:
        j1 = 2 * i + 1
        if j1 >= n or j1 < 0:
            break
        j = j1
        j2 = j1 + 1
        if j2 < n and h.Less(j2, j1):
            j = j2
        if not h.Less(j, i):
            break
        h.Swap(i, j)
        i = j
    return i > i0


def heap_Init(h):
    n = h.Len()
    for i in range(n // 2 - 1, -1, -1):
        _heap_down(h, i, n)


def heap_Push(h, x):
    h.Push(x)
    _heap_up(h, h.Len() - 1)


def heap_Pop(h):
    n = h.Len() - 1
    if n < 0:
        raise GoPanic("index out of range [-1]")
    h.Swap(0, n)
    _heap_down(h, 0, n)
    return h.Pop()


def buildTree(symFreqs: dict):  # huffman.go:58-103
    keys = []
    values = []
    for i, j in symFreqs.items():
        keys.append(int(i))
        values.append(j)
    keys.sort()
    values.sort()
    temp1, temp2 = [], []
    for value in list(values):
        for i, key in enumerate(keys):
            if symFreqs[key] == value:
                temp1.append(key)
                temp2.append(value)
                keys[i] = keys[len(keys) - 1]  # remove(): swap-with-last
                keys = keys[: len(keys) - 1]
                keys.sort()
                values.sort()
                break
    trees = treeHeap()
    for i in range(len(symFreqs)):
        trees.append(HuffmanLeaf(temp2[i], temp1[i]))
    heap_Init(trees)
    while trees.Len() > 1:
        a = heap_Pop(trees)
        b = heap_Pop(trees)
        heap_Push(trees, HuffmanNode(a.Freq() + b.Freq(), a, b))
    return heap_Pop(trees)


def printCodes(tree, prefix: bytearray, vals: list, bins: list):  # huffman.go:110-127
    if isinstance(tree, HuffmanLeaf):
        vals.append(tree.value)
        bins.append(bytes(prefix).decode())
        return vals, bins
    prefix.append(0x30)
    printCodes(tree.left, prefix, vals, bins)
    prefix.pop()
    prefix.append(0x31)
    printCodes(tree.right, prefix, vals, bins)
    prefix.pop()
    return vals, bins


def AsByteSlice(b: str) -> bytes:  # huffman.go:174-191
    out = bytearray()
    i = len(b)
    while i > 0:
        s = b[0:i] if i - 8 < 0 else b[i - 8 : i]
        out[0:0] = bytes([int(s, 2)])
        i -= 8
    return bytes(out)


def huff_encode(tree, inp: bytes, estring: bytes) -> bytes:  # huffman.go:229-256
    vals, bins = printCodes(tree, bytearray(), [], [])
    answer = []
    for _, c in go_range_string(inp):
        if c in vals:
            answer.append(bins[vals.index(c)])
        else:
            answer.append(bins[0])
    answer = "".join(answer)
    diff = format(8 - len(answer) % 8, "b")
    if diff == "1000":
        diff = "0"
    first = AsByteSlice(diff)
    final = AsByteSlice(answer)
    return estring + b"\\\n" + first + final


def huff_Compress(fileContents: bytes, order=None) -> bytes:  # huffman.go:299-325
    """`order`: optional function mapping the list of runes to the header record order
    (Go ranges over a map, i.e. unspecified order).  Default: insertion (first-seen) order."""
    symFreqs = {}
    for _, c in go_range_string(fileContents):
        symFreqs[c] = symFreqs.get(c, 0) + 1
    ks = list(symFreqs.keys())
    if order is not None:
        ks = order(ks)
    estring = bytearray()
    for key in ks:
        val = symFreqs[key]
        if key != 10:
            estring += str(val).encode() + b"|" + go_string_rune(key)
        else:
            estring += str(val).encode() + b"|\\n"
    exampleTree = buildTree(symFreqs)
    return huff_encode(exampleTree, fileContents, bytes(estring))


def decodeTree(tree: bytes):  # huffman.go:196-227
    symFreqs = {}
    temp = bytearray()
    i = 0
    n = len(tree)
    while i < n:
        if tree[i] != 0x7C:
            if 0x30 <= tree[i] <= 0x39:
                temp.append(tree[i])
        else:
            freq = go_atoi(bytes(temp).strip())
            temp = bytearray()
            if i + 1 >= n:
                raise GoPanic("index out of range")
            if tree[i + 1] == 0x5C:
                if i + 2 >= n:
                    raise GoPanic("index out of range")
            if tree[i + 1] == 0x5C and tree[i + 2] == 0x6E:
                symFreqs[10] = freq
                i += 1
            else:
                # Go: `for j, c := range tree { if j == i+1 {...; break} }` rescans from 0.
                # tree[i] is ASCII '|', so i+1 is always a rune boundary of that scan and the
                # rune found there equals the one decoded directly at i+1.
                for j, c in go_range_string(tree[i + 1 : i + 5]):
                    symFreqs[c] = freq
                    break
            i += 1
        i += 1
    return symFreqs


def findCodes(tree, og, data: str, i: int, maxi: int, strict: bool, answer: bytearray):  # huffman.go:131-153
    # tail recursion unrolled into a loop; one iteration per Go call
    while True:
        if strict and i > 900000:
            raise GoPanic("Max recursion depth")
        if i <= maxi:
            if isinstance(tree, HuffmanLeaf):
                answer += go_string_rune(tree.value)
                if i < maxi:
                    if tree is og:
                        raise GoPanic("unbounded recursion (single-leaf tree with bits left)")
                    tree = og
                    continue
                return bytes(answer)
            if i >= len(data):
                raise GoPanic("index out of range")
            if data[i] == "0":
                tree = tree.left
            else:
                tree = tree.right
            i += 1
            continue
        return bytes(answer)


def huff_Decompress(fileContents: bytes, strict: bool = False) -> bytes:  # huffman.go:258-297
    k = fileContents.find(b"\\\n")
    if k < 0:
        raise GoPanic("index out of range [1] with length 1")
    sections = [fileContents[:k], fileContents[k + 2 :]]
    symFreqs = decodeTree(sections[0])
    tree = buildTree(symFreqs)
    byteArr = sections[1]
    content = []
    diff = 0
    for i, nb in enumerate(byteArr):
        if i != 0:
            content.append(format(nb, "08b"))
        else:
            diff = nb
    contentString = "".join(content)
    if diff > len(contentString):
        raise GoPanic("slice bounds out of range")
    data = contentString[diff:]
    return findCodes(tree, tree, data, 0, len(data), strict, bytearray())
