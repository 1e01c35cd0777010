// Seam A: IStreamDecoder on the reference's OWN packet provider (Ogg layer, NewStream plumbing, stats of the
// container stay untouched); header parse, entropy decode, floor, residue, coupling, IMDCT and overlap-add run on
// the GPU through vpz_setup_* / vpz_batch_*.  VorbisReader.cs:70 becomes `new GpuStreamDecoder(packetProvider, gpu)`.
//
// The stream is decoded AHEAD in windows of `Lookahead` packets: ReadNextPacket's position / end-of-stream
// bookkeeping (StreamDecoder.cs:640-694) depends only on packet headers, so it is replayed while a window is planned
// and the GPU decodes the whole window in one pass.  Read then hands the cached PCM out with the reference's
// call-by-call behaviour (at most one packet per call).  This is the C# rendering of csrc/reader.cpp
// (plan_window / resolve_drain / decode_ahead / read_next_packet / stream_read / seek_begin / seek_finish), which is
// what the parity tests run against the oracle.
using System;
using System.Collections.Generic;
using System.IO;
using NVorbis.Contracts;

namespace NVorbis.Gpu
{
    public sealed unsafe class GpuStreamDecoder : IStreamDecoder, IPacketGranuleCountProvider
    {
        [Flags] enum Eos { None = 0, InvalidPacket = 1, PacketFlag = 2, InvalidPreroll = 4 }   // EndOfStreamFlags.cs

        struct Entry              // one packet handed out by the provider, already planned / decoded
        {
            public bool Ok, IsResync;
            public Eos EosFlags;
            public Exception? Error;   // what the reference throws at this packet
            public int Count, Tail;    // samples made available; overlap length left for the next packet
            public long Granule;
            public int PcmOffset, DrainOffset;   // float offsets in _pcm
        }

        readonly IPacketProvider _packets;
        readonly GpuContext _gpu;
        IntPtr _setup, _batch;
        VpzSetupInfo _info;
        byte[] _vendor = Array.Empty<byte>();
        byte[][] _comments = Array.Empty<byte[]>();
        ITagData? _tags;

        // consumption side (StreamDecoder.cs:40-49)
        bool _havePrev, _hasPosition, _hasClipped, _disposed;
        int _prevAvail, _prevTail, _pcmCur;
        Eos _eosFound;
        long _currentPosition;
        // decode-ahead state
        readonly List<Entry> _queue = new();
        int _queueHead;
        float[] _pcm = Array.Empty<float>();
        byte[]? _carry;            // last decoded packet: re-submitted as the seed of the next window
        int _carryTrim;
        int _seekLeft;             // packets SeekTo still has to consume (they see the stale position)
        long _seekPos;

        public int Lookahead { get; set; } = 256;
        public bool ClipSamples { get; set; } = true;
        public bool SkipTags { get; set; }

        public GpuStreamDecoder(IPacketProvider packetProvider, GpuContext? gpu = null)
        {
            _packets = packetProvider ?? throw new ArgumentNullException(nameof(packetProvider));
            _gpu = gpu ?? GpuContext.Default;
        }

        IntPtr Ctx => _gpu.Handle;

        static byte[] Bytes(ref VorbisPacket p)
        {
            byte[] b = new byte[(p.TotalBits + 7) / 8];
            p.Reset();
            p.ReadBytes(b);
            return b;
        }

        // ---- StreamDecoder.Initialize / ProcessHeaderPackets (StreamDecoder.cs:71-184) ------------------------
        public void Initialize()
        {
            byte[]? id = null, books = null;
            for (int want = 1; want <= 5; want += 2)
            {
                VorbisPacket p = _packets.GetNextPacket();
                if (!p.IsValid) throw new InvalidDataException("Could not find Vorbis data to decode.");
                byte[] b = Bytes(ref p);
                p.Finish();
                if (b.Length < 7 || b[0] != want || b[1] != 'v' || b[2] != 'o' || b[3] != 'r' || b[4] != 'b' || b[5] != 'i' || b[6] != 's')
                    throw new InvalidDataException("Vorbis header packets are out of order or missing.");
                if (want == 1) id = b;
                else if (want == 3) LoadComments(b);
                else books = b;
            }
            fixed (byte* a = id, c = books)
                Vpz.Check(Vpz.vpz_setup_create(Ctx, a, (nuint)id!.Length, c, (nuint)books!.Length, out _setup), Ctx);
            Vpz.Check(Vpz.vpz_setup_get_info(_setup, out _info), Ctx);
            Vpz.Check(Vpz.vpz_batch_create(Ctx, out _batch), Ctx);
            ResetDecoder();
        }

        void LoadComments(byte[] b)                                  // StreamDecoder.cs:242-260
        {
            if (SkipTags) return;
            int at = 7;
            byte[] Str()
            {
                if (at + 4 > b.Length) return Array.Empty<byte>();
                int n = BitConverter.ToInt32(b, at);
                at += 4;
                n = Math.Max(0, Math.Min(n, b.Length - at));
                byte[] s = b.AsSpan(at, n).ToArray();
                at += n;
                return s;
            }
            _vendor = Str();
            int count = at + 4 <= b.Length ? BitConverter.ToInt32(b, at) : 0;
            at += 4;
            List<byte[]> list = new();
            for (int i = 0; i < count && at < b.Length; i++) list.Add(Str());
            _comments = list.ToArray();
        }

        void ResetDecoder()                                          // StreamDecoder.cs:357-369
        {
            _havePrev = false;
            _prevAvail = _prevTail = 0;
            _eosFound = Eos.None;
            _hasClipped = false;
            _hasPosition = false;
            _queue.Clear();
            _queueHead = 0;
            _carry = null;
            _carryTrim = 0;
        }

        // ---- IPacketGranuleCountProvider (StreamDecoder.cs:882-913) ---------------------------------------------
        int IPacketGranuleCountProvider.GetPacketGranuleCount(ref VorbisPacket p)
        {
            if (p.IsResync) return 0;
            byte[] d = Bytes(ref p);
            int* info = stackalloc int[6];
            fixed (byte* q = d)
                return Vpz.vpz_packet_info(_setup, q, (nuint)d.Length, info) == 1 ? info[4] - info[2] : 0;
        }

        // ---- window planning: reader.cpp plan_window + resolve_drain ---------------------------------------------
        sealed class Window
        {
            public readonly List<Entry> Entries = new();
            public readonly List<byte[]> Src = new();       // submitted packets
            public readonly List<int> Trims = new();
            public readonly List<int> EntryOf = new();       // submitted packet -> entry (-1: the carried-over seed)
            public bool Drain, DrainExtraRun;
            public int DrainTail, DrainEntry = -1;
        }

        Window PlanWindow(int maxPackets)
        {
            Window w = new();
            bool havePrev = _havePrev, hasPos = _hasPosition;
            int prevTail = _prevTail, seekLeft = _seekLeft;
            long pos = _currentPosition;
            if (havePrev && _carry != null)
            {
                w.Src.Add(_carry);
                w.Trims.Add(_carryTrim);
                w.EntryOf.Add(-1);
            }
            int* info = stackalloc int[6];
            for (int n = 0; maxPackets <= 0 || n < maxPackets; n++)
            {
                VorbisPacket pk = _packets.GetNextPacket();
                Entry e = default;
                e.Granule = -1;
                if (!pk.IsValid)                                        // StreamDecoder.cs:703-711
                {
                    e.EosFlags = Eos.InvalidPacket;
                    w.Entries.Add(e);
                    break;
                }
                byte[] data = Bytes(ref pk);
                e.EosFlags = pk.IsEndOfStream ? Eos.PacketFlag : Eos.None;
                e.IsResync = pk.IsResync;
                long granule = pk.GranulePosition;
                pk.Finish();
                if (e.IsResync) hasPos = false;
                int rc;
                fixed (byte* q = data) rc = Vpz.vpz_packet_info(_setup, q, (nuint)data.Length, info);
                if (rc == Vpz.InvalidData)                              // unused mode: the exception leaves DecodeNextPacket
                {                                                       // before _eosFound is updated (StreamDecoder.cs:732-735)
                    e.EosFlags = Eos.None;
                    e.Error = new InvalidDataException("Invalid mode index.");
                    w.Entries.Add(e);
                    continue;
                }
                if (rc != 1)                                            // not an audio packet: yields nothing
                {
                    w.Entries.Add(e);
                    if (e.EosFlags != Eos.None || seekLeft > 0) break;  // a failed packet ends SeekTo early
                    continue;
                }
                int length = prevTail, leftUse1 = info[1], leftStart = info[2], rightStart = info[4], rightEnd = info[5];
                int trim = 0;
                if (granule != -1 && e.EosFlags != Eos.None)            // StreamDecoder.cs:658-666
                {
                    int diff = (int)(pos + length - granule);
                    if (diff > 0)
                    {
                        trim = diff;
                        rightStart = Math.Max(rightStart - diff, 0);
                    }
                }
                if (havePrev)
                {
                    int slope = (leftUse1 != 0 ? _info.BlockSize1 : _info.BlockSize0) / 2;
                    if (length > slope || length < 0 || leftStart + length > _info.BlockSize1 || rightStart < leftStart)
                    {
                        e.Error = new ArgumentOutOfRangeException("OverlapBuffers");   // the reference faults here (SURVEY quirk Q4)
                        w.Entries.Add(e);
                        break;
                    }
                }
                e.Ok = true;
                e.Count = havePrev ? rightStart - leftStart : 0;
                e.Tail = rightEnd - rightStart;
                e.Granule = granule;
                w.Src.Add(data);
                w.Trims.Add(trim);
                w.EntryOf.Add(w.Entries.Count);
                w.Entries.Add(e);
                havePrev = true;
                prevTail = e.Tail;
                if (seekLeft > 0)
                {
                    if (--seekLeft == 0) pos = _seekPos + e.Count;      // after SeekTo's target packet
                }
                else
                {
                    if (granule != -1 && !hasPos)                       // StreamDecoder.cs:459-463
                    {
                        hasPos = true;
                        pos = granule - e.Count;
                    }
                    pos += e.Count;
                }
                if (e.EosFlags != Eos.None) break;
            }
            // StreamDecoder.cs:451-455: an end-of-stream packet that fails to decode drains the previous packet's raw right half
            if (w.Entries.Count > 0 && w.EntryOf.Count > 0)
            {
                Entry last = w.Entries[^1];
                if (!last.Ok && last.Error == null && (last.EosFlags & Eos.PacketFlag) != 0)
                {
                    int k = w.EntryOf.Count - 1, ei = w.EntryOf[k];
                    int tail = ei >= 0 ? w.Entries[ei].Tail : _prevTail;
                    if (tail > 0)
                    {
                        w.Drain = true;
                        w.DrainTail = tail;
                        w.DrainEntry = w.Entries.Count - 1;
                        if (ei >= 0) w.Trims[k] = -tail; else w.DrainExtraRun = true;
                    }
                }
            }
            return w;
        }

        int AddRun(List<byte[]> src, List<int> trims)
        {
            int total = 0;
            foreach (byte[] s in src) total += s.Length;
            byte[] flat = new byte[Math.Max(total, 1)];
            uint[] off = new uint[src.Count + 1];
            int at = 0;
            for (int i = 0; i < src.Count; i++)
            {
                off[i] = (uint)at;
                src[i].CopyTo(flat, at);
                at += src[i].Length;
            }
            off[src.Count] = (uint)at;
            int[] tr = trims.ToArray();
            fixed (byte* b = flat)
            fixed (uint* o = off)
            fixed (int* t = tr)
            {
                int run = Vpz.vpz_batch_add_run(_batch, _setup, b, o, (uint)src.Count, t);
                Vpz.Check(run, Ctx);
                return run;
            }
        }

        // reader.cpp decode_ahead + place_window
        void DecodeAhead()
        {
            Window w = PlanWindow(Lookahead);
            Vpz.Check(Vpz.vpz_batch_reset(_batch), Ctx);
            int run = -1, drainRun = -1;
            bool emits = false;
            foreach (int e in w.EntryOf) emits |= e >= 0;
            if (w.Src.Count > 0 && emits) run = AddRun(w.Src, w.Trims);
            if (w.DrainExtraRun)
            {
                byte[] p = w.Src[^1];
                drainRun = AddRun(new List<byte[]> { p, p }, new List<int> { 0, -w.DrainTail });
            }
            int total = (int)Vpz.vpz_batch_total_floats(_batch);
            if (total > 0)
            {
                if (_pcm.Length < total) _pcm = new float[total];
                Vpz.Check(Vpz.vpz_batch_decode(_batch, 0), Ctx);   // unclipped: Read clips while it copies (per-sample HasClipped)
                fixed (float* d = _pcm) Vpz.Check(Vpz.vpz_batch_read_all(_batch, d), Ctx);
            }
            int channels = _info.Channels;
            if (run >= 0)
            {
                int[] counts = new int[w.Src.Count];
                fixed (int* c = counts) Vpz.Check(Vpz.vpz_batch_run_packet_samples(_batch, run, c), Ctx);
                int status = Vpz.vpz_batch_run_status(_batch, run, out int stop);
                int off = (int)Vpz.vpz_batch_run_offset(_batch, run);
                for (int k = 0; k < w.EntryOf.Count; k++)
                {
                    int ei = w.EntryOf[k];
                    if (ei >= 0)
                    {
                        Entry e = w.Entries[ei];
                        if (status == Vpz.RefFault && k >= stop)
                        {
                            e.Ok = false;
                            e.Error = new ArgumentOutOfRangeException("OverlapBuffers");
                        }
                        else
                        {
                            e.PcmOffset = off;
                            if (w.Drain && !w.DrainExtraRun && k + 1 == w.EntryOf.Count)
                            {
                                Entry de = w.Entries[w.DrainEntry];
                                de.DrainOffset = off + e.Count * channels;
                                w.Entries[w.DrainEntry] = de;
                            }
                        }
                        w.Entries[ei] = e;
                    }
                    off += counts[k] * channels;
                }
            }
            if (drainRun >= 0)
            {
                Entry de = w.Entries[w.DrainEntry];
                de.DrainOffset = (int)Vpz.vpz_batch_run_offset(_batch, drainRun) + (int)(Vpz.vpz_batch_run_samples(_batch, drainRun) - w.DrainTail) * channels;
                w.Entries[w.DrainEntry] = de;
            }
            for (int k = w.EntryOf.Count - 1; k >= 0; k--)      // the last decoded packet seeds the next window
            {
                if (w.EntryOf[k] < 0) break;
                if (w.Entries[w.EntryOf[k]].Ok)
                {
                    _carry = w.Src[k];
                    _carryTrim = w.Trims[k];
                    break;
                }
            }
            _queue.Clear();
            _queue.AddRange(w.Entries);
            _queueHead = 0;
        }

        // ---- StreamDecoder.ReadNextPacket (StreamDecoder.cs:640-694), consumption side --------------------------
        bool ReadNextPacket(out long samplePosition)
        {
            samplePosition = -1;
            if (_queueHead == _queue.Count)
            {
                DecodeAhead();
                if (_queue.Count == 0)
                {
                    _eosFound |= Eos.InvalidPacket;
                    return false;
                }
            }
            Entry e = _queue[_queueHead++];
            if (_seekLeft > 0) _seekLeft--;
            if (e.IsResync) _hasPosition = false;
            if (e.Error != null) throw e.Error;
            _eosFound |= e.EosFlags;
            if (!e.Ok)
            {
                if ((e.EosFlags & Eos.PacketFlag) != 0 && _havePrev) _pcmCur = e.DrainOffset;
                return false;
            }
            samplePosition = e.Granule;
            _prevAvail = _havePrev ? e.Count : 0;
            _prevTail = e.Tail;
            _pcmCur = e.PcmOffset;
            _havePrev = true;
            return true;
        }

        // ---- StreamDecoder.Read (StreamDecoder.cs:407-498) --------------------------------------------------------
        public int Read(Span<float> buffer) => Read(buffer, buffer.Length / Math.Max(_info.Channels, 1), 0, true);
        public int Read(Span<float> buffer, int samplesToRead, int channelStride) => Read(buffer, samplesToRead, channelStride, false);

        int Read(Span<float> buffer, int samplesToRead, int channelStride, bool interleave)
        {
            if (_disposed) throw new ObjectDisposedException(nameof(GpuStreamDecoder));
            int channels = _info.Channels;
            if (buffer.Length % channels != 0) throw new ArgumentException("Length must be a multiple of Channels.", nameof(buffer));
            if (buffer.Length < samplesToRead * channels) throw new ArgumentException("The buffer is too small.", nameof(buffer));
            int idx = 0;
            while (idx == 0)
            {
                if (_prevAvail == 0)
                {
                    if (_eosFound != Eos.None)
                    {
                        _havePrev = false;
                        break;
                    }
                    if (!ReadNextPacket(out long sp))
                    {
                        if ((_eosFound & Eos.PacketFlag) != 0)             // StreamDecoder.cs:451-455
                        {
                            _prevAvail = _prevTail;
                            _prevTail = 0;
                        }
                    }
                    else if (sp != -1 && !_hasPosition)
                    {
                        _hasPosition = true;
                        _currentPosition = sp - _prevAvail - idx;
                    }
                }
                int copy = Math.Min(samplesToRead - idx, _prevAvail);
                if (copy <= 0)
                {
                    if (samplesToRead - idx <= 0) break;
                    continue;
                }
                ReadOnlySpan<float> src = _pcm.AsSpan(_pcmCur, copy * channels);
                if (interleave)
                {
                    Span<float> dst = buffer.Slice(idx * channels, copy * channels);
                    if (ClipSamples) for (int i = 0; i < src.Length; i++) dst[i] = Utils.ClipValue(src[i], ref _hasClipped);
                    else src.CopyTo(dst);
                }
                else
                {
                    for (int ch = 0; ch < channels; ch++)
                        for (int i = 0; i < copy; i++)
                        {
                            float v = src[i * channels + ch];
                            buffer[ch * channelStride + idx + i] = ClipSamples ? Utils.ClipValue(v, ref _hasClipped) : v;
                        }
                }
                idx += copy;
                _pcmCur += copy * channels;
                _prevAvail -= copy;
                _currentPosition += copy;
            }
            return idx;
        }

        // ---- StreamDecoder.SeekTo (StreamDecoder.cs:803-880) -------------------------------------------------------
        public void SeekTo(TimeSpan timePosition, SeekOrigin seekOrigin = SeekOrigin.Begin)
            => SeekTo((long)(_info.SampleRate * timePosition.TotalSeconds), seekOrigin);

        public void SeekTo(long samplePosition, SeekOrigin seekOrigin = SeekOrigin.Begin)
        {
            if (_disposed) throw new ObjectDisposedException(nameof(GpuStreamDecoder));
            if (!_packets.CanSeek) throw new InvalidOperationException("Seek is not supported by the underlying packet provider.");
            if (samplePosition < 0) throw new ArgumentOutOfRangeException(nameof(samplePosition));
            switch (seekOrigin)
            {
                case SeekOrigin.Begin: break;
                case SeekOrigin.Current: samplePosition = SamplePosition - samplePosition; break;
                case SeekOrigin.End: samplePosition = TotalSamples - samplePosition; break;
                default: throw new ArgumentOutOfRangeException(nameof(seekOrigin));
            }
            long pos = _packets.SeekTo(samplePosition, 1, this);        // pre-roll of one packet, the reference's own code
            int rollForward = (int)(samplePosition - pos);
            ResetDecoder();
            _hasPosition = true;
            _seekLeft = 2;                                               // the window planner sees the stale position for these two
            _seekPos = pos;
            try
            {
                if (!ReadNextPacket(out _))
                {
                    _eosFound |= Eos.InvalidPreroll;
                    if (samplePosition > _packets.GetGranuleCount(this)) throw new SeekOutOfRangeException();
                    _prevAvail = 0;
                    _currentPosition = samplePosition;
                    return;
                }
                if (!ReadNextPacket(out _))
                {
                    ResetDecoder();
                    _eosFound |= Eos.InvalidPacket;
                    throw new PreRollPacketException();
                }
                _prevAvail -= rollForward;
                _pcmCur += rollForward * _info.Channels;
                _currentPosition = samplePosition;
            }
            finally
            {
                _seekLeft = 0;
            }
        }

        // ---- the rest of IStreamDecoder (StreamDecoder.cs:933-1007) -------------------------------------------------
        public int StreamSerial => _packets.StreamSerial;
        public int Channels => _info.Channels;
        public int SampleRate => _info.SampleRate;
        public int UpperBitrate => _info.BitrateUpper;
        public int NominalBitrate => _info.BitrateNominal;
        public int LowerBitrate => _info.BitrateLower;
        public ITagData Tags => _tags ??= new TagData(_vendor, _comments);
        public long TotalSamples => _packets.GetGranuleCount(this);
        public TimeSpan TotalTime => TimeSpan.FromSeconds((double)TotalSamples / _info.SampleRate);
        public bool HasClipped => _hasClipped;
        public bool IsEndOfStream => _eosFound != Eos.None && _prevAvail == 0;     // StreamDecoder.cs:1001
        public IStreamStats Stats => throw new NotSupportedException("stats counters are not kept by the GPU path");
        public long SamplePosition { get => _currentPosition; set => SeekTo(value); }
        public TimeSpan TimePosition
        {
            get => TimeSpan.FromSeconds((double)_currentPosition / _info.SampleRate);
            set => SeekTo(value);
        }

        public void Dispose()
        {
            if (_disposed) return;
            _disposed = true;
            if (_batch != IntPtr.Zero) Vpz.vpz_batch_destroy(_batch);
            if (_setup != IntPtr.Zero) Vpz.vpz_setup_release(_setup);
            _batch = _setup = IntPtr.Zero;
            _packets.Dispose();
        }
    }
}
