// Seam B: the public reader surface forwarded to vpz_reader_* (our own Ogg layer + GPU decode).
// Member for member the class vorbispizza_b200/api.py:VorbisReader that tests/cases.py drives against the oracle.
using System;
using System.Collections.Generic;
using System.IO;
using System.Runtime.InteropServices;
using NVorbis.Contracts;

namespace NVorbis.Gpu
{
    public sealed unsafe class GpuVorbisReader : IVorbisReader
    {
        readonly GpuContext _gpu;
        readonly byte[] _image;              // container bytes; pinned for the reader's lifetime (copy = 0)
        GCHandle _pin;
        IntPtr _reader;
        readonly List<IStreamDecoder> _streams = new();
        ITagData? _tags;
        int _tagsOfStream = -1;

        public event NewStreamEventHandler? NewStream;

        public GpuVorbisReader(Stream stream, GpuContext? gpu = null) : this(ReadAll(stream), gpu) { }

        public GpuVorbisReader(byte[] containerImage, GpuContext? gpu = null)
        {
            _gpu = gpu ?? GpuContext.Default;
            _image = containerImage;
            _pin = GCHandle.Alloc(_image, GCHandleType.Pinned);
        }

        static byte[] ReadAll(Stream s)
        {
            using MemoryStream ms = new();
            s.CopyTo(ms);
            return ms.ToArray();
        }

        IntPtr Ctx => _gpu.Handle;
        IntPtr R => _reader != IntPtr.Zero ? _reader : throw new ObjectDisposedException(nameof(GpuVorbisReader));

        /// <summary>VorbisReader.Initialize (VorbisReader.cs:92-103): opens the first logical stream.</summary>
        public void Initialize()
        {
            int rc = Vpz.vpz_reader_open_memory(Ctx, (byte*)_pin.AddrOfPinnedObject(), (nuint)_image.Length, 0, out _reader);
            Vpz.Check(rc, Ctx);
            RaiseNewStreams(0);
        }

        // The ABI has no callbacks: NewStream is raised here whenever the native stream count has grown.
        void RaiseNewStreams(int before)
        {
            int now = Vpz.vpz_reader_stream_count(R);
            for (int i = before; i < now; i++)
            {
                GpuLogicalStream dec = new(this, i);
                _streams.Add(dec);
                NewStreamEventArgs ea = new(dec);
                NewStream?.Invoke(this, ref ea);
                // IgnoreStream: the managed list simply hides the stream (VorbisReader.cs:72-84); the native reader keeps it
                if (ea.IgnoreStream) _streams.Remove(dec);
            }
        }

        public bool FindNextStream()                               // VorbisReader.cs:183-189
        {
            int before = Vpz.vpz_reader_stream_count(R);
            int rc = Vpz.vpz_reader_find_next_stream(R);
            Vpz.Check(rc, Ctx);
            if (rc == 1) RaiseNewStreams(before);
            return rc == 1;
        }

        public bool SwitchStreams(int index)                       // VorbisReader.cs:191-210
        {
            if (index < 0 || index >= Vpz.vpz_reader_stream_count(R)) throw new ArgumentOutOfRangeException(nameof(index));
            int rc = Vpz.vpz_reader_switch_stream(R, index);
            Vpz.Check(rc, Ctx);
            return rc == 1;                                        // 1: channel count or sample rate changed
        }

        public IReadOnlyList<IStreamDecoder> Streams => _streams;
        public int StreamIndex => Vpz.vpz_reader_stream_index(R);
        public bool CanSeek => Vpz.vpz_reader_can_seek(R) == 1;
        public long ContainerOverheadBits => Vpz.vpz_reader_container_overhead_bits(R);
        public long ContainerWasteBits => Vpz.vpz_reader_container_waste_bits(R);
        public int StreamSerial => Vpz.vpz_reader_stream_serial(R);
        public int Channels => Vpz.vpz_reader_channels(R);
        public int SampleRate => Vpz.vpz_reader_sample_rate(R);
        public int UpperBitrate => Vpz.vpz_reader_bitrate(R, 0);
        public int NominalBitrate => Vpz.vpz_reader_bitrate(R, 1);
        public int LowerBitrate => Vpz.vpz_reader_bitrate(R, 2);
        public long TotalSamples { get { long n = Vpz.vpz_reader_total_samples(R); Vpz.Check(n, Ctx); return n; } }
        public TimeSpan TotalTime => TimeSpan.FromSeconds((double)TotalSamples / SampleRate);
        public bool HasClipped => Vpz.vpz_reader_has_clipped(R) == 1;
        public bool IsEndOfStream => Vpz.vpz_reader_is_end_of_stream(R) == 1;
        public IStreamStats StreamStats => throw new NotSupportedException("stats counters stay with the managed decoder (out of scope of the GPU path)");

        public bool ClipSamples
        {
            get => Vpz.vpz_reader_get_clip(R) == 1;
            set => Vpz.vpz_reader_set_clip(R, value ? 1 : 0);
        }

        public long SamplePosition
        {
            get => Vpz.vpz_reader_sample_position(R);
            set => SeekTo(value);
        }

        public TimeSpan TimePosition
        {
            get => TimeSpan.FromSeconds((double)SamplePosition / SampleRate);
            set => SeekTo(value);
        }

        public ITagData Tags
        {
            get
            {
                int cur = StreamIndex;
                if (_tags == null || _tagsOfStream != cur)
                {
                    byte[] vendor = Bytes(Vpz.vpz_reader_vendor(R, out int vl), vl);
                    int n = Vpz.vpz_reader_comment_count(R);
                    byte[][] comments = new byte[n][];
                    for (int i = 0; i < n; i++) comments[i] = Bytes(Vpz.vpz_reader_comment(R, i, out int cl), cl);
                    _tags = new TagData(vendor, comments);          // TagData.cs: the reference's own parser of the raw strings
                    _tagsOfStream = cur;
                }
                return _tags;
            }
        }

        static byte[] Bytes(IntPtr p, int len)
        {
            byte[] b = new byte[Math.Max(len, 0)];
            if (len > 0) Marshal.Copy(p, b, 0, len);
            return b;
        }

        public int ReadSamples(Span<float> buffer)                 // VorbisReader.cs:232-241
        {
            int count = buffer.Length - buffer.Length % Channels;
            if (count == 0) return 0;
            fixed (float* p = buffer)
            {
                int n = Vpz.vpz_reader_read(R, p, count);
                Vpz.Check(n, Ctx);
                return n;
            }
        }

        public int ReadSamples(Span<float> buffer, int samplesToRead, int channelStride)   // VorbisReader.cs:243-252
        {
            fixed (float* p = buffer)
            {
                int n = Vpz.vpz_reader_read_planar(R, p, buffer.Length, samplesToRead, channelStride);
                Vpz.Check(n, Ctx);
                return n;
            }
        }

        public void SeekTo(long samplePosition, SeekOrigin seekOrigin = SeekOrigin.Begin)   // VorbisReader.cs:226
            => Vpz.Check(Vpz.vpz_reader_seek(R, samplePosition, (int)seekOrigin), Ctx);

        public void SeekTo(TimeSpan timePosition, SeekOrigin seekOrigin = SeekOrigin.Begin) // VorbisReader.cs:220
            => SeekTo((long)(SampleRate * timePosition.TotalSeconds), seekOrigin);

        /// <summary>Packets decoded ahead per GPU pass (default 256).</summary>
        public int Lookahead { set => Vpz.Check(Vpz.vpz_reader_set_lookahead(R, value), Ctx); }

        public void Dispose()
        {
            if (_reader != IntPtr.Zero) Vpz.vpz_reader_close(_reader);
            _reader = IntPtr.Zero;
            if (_pin.IsAllocated) _pin.Free();
        }

        /// <summary>One logical stream as IStreamDecoder: switches the native reader to it, then forwards.</summary>
        sealed class GpuLogicalStream : IStreamDecoder
        {
            readonly GpuVorbisReader _o;
            readonly int _index;
            internal GpuLogicalStream(GpuVorbisReader owner, int index) { _o = owner; _index = index; }
            GpuVorbisReader On() { if (_o.StreamIndex != _index) _o.SwitchStreams(_index); return _o; }
            public int StreamSerial => On().StreamSerial;
            public int Channels => On().Channels;
            public int SampleRate => On().SampleRate;
            public int UpperBitrate => On().UpperBitrate;
            public int NominalBitrate => On().NominalBitrate;
            public int LowerBitrate => On().LowerBitrate;
            public ITagData Tags => On().Tags;
            public TimeSpan TotalTime => On().TotalTime;
            public long TotalSamples => On().TotalSamples;
            public TimeSpan TimePosition { get => On().TimePosition; set => On().TimePosition = value; }
            public long SamplePosition { get => On().SamplePosition; set => On().SamplePosition = value; }
            public bool ClipSamples { get => On().ClipSamples; set => On().ClipSamples = value; }
            public bool SkipTags { get; set; }
            public bool HasClipped => On().HasClipped;
            public bool IsEndOfStream => On().IsEndOfStream;
            public IStreamStats Stats => On().StreamStats;
            public void Initialize() { }
            public void SeekTo(TimeSpan t, SeekOrigin o = SeekOrigin.Begin) => On().SeekTo(t, o);
            public void SeekTo(long s, SeekOrigin o = SeekOrigin.Begin) => On().SeekTo(s, o);
            public int Read(Span<float> buffer) => On().ReadSamples(buffer);
            public int Read(Span<float> buffer, int samplesToRead, int channelStride) => On().ReadSamples(buffer, samplesToRead, channelStride);
            public void Dispose() { }
        }
    }
}
